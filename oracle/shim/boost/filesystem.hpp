// ORACLE/shim: nothing of boost::filesystem is used by the compiled reference headers
