// ORACLE/shim: placeholder so that the reference header that names this file can be parsed; nothing of it is used on the compiled path
