/* tests/stubs/ruby/config.h — TEST SCAFFOLDING (see ruby.h) */
