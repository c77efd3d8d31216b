/* tests/stubs/ruby/ruby.h — TEST SCAFFOLDING: the slice of the Ruby C API that the reference's
 * src/smatrix_ruby.c uses, enough to COMPILE AND LINK it unchanged against include/smatrix.h and
 * our library (no Ruby in this image).  Not a working interpreter interface. */
#ifndef SMX_STUB_RUBY_H
#define SMX_STUB_RUBY_H
#include <stdint.h>
typedef uintptr_t VALUE;
typedef uintptr_t ID;
enum { RUBY_T_NIL = 0x11, RUBY_T_DATA = 0x0c, RUBY_T_STRING = 0x05, RUBY_T_FIXNUM = 0x15 };
#define T_STRING RUBY_T_STRING
#define Qnil ((VALUE)8)
extern VALUE rb_cObject, rb_eTypeError;
VALUE rb_iv_get(VALUE, const char*);
VALUE rb_iv_set(VALUE, const char*, VALUE);
int rb_type(VALUE);
void rb_raise(VALUE, const char*, ...);
VALUE rb_define_class(const char*, VALUE);
void rb_define_method(VALUE, const char*, VALUE (*)(), int);
char* smx_stub_rstring_ptr(VALUE);
void* smx_stub_data_ptr(VALUE);
VALUE smx_stub_data_wrap(VALUE, void*, void*, void*);
long smx_stub_num2int(VALUE);
VALUE smx_stub_int2num(long);
#define RSTRING_PTR(v) smx_stub_rstring_ptr(v)
#define NUM2INT(v) ((int)smx_stub_num2int(v))
#define INT2NUM(v) smx_stub_int2num((long)(v))
#define Data_Get_Struct(obj, type, sval) ((sval) = (type*)smx_stub_data_ptr(obj))
#define Data_Wrap_Struct(klass, mark, free, sval) smx_stub_data_wrap((klass), (void*)(mark), (void*)(free), (sval))
#endif
