// TEST INFRASTRUCTURE.  Configuration header for building the reference's vendored fparser 4.5.2 where it lies
// (/root/reference/3rdparty/fparser-4.5.2/fparser.cc) into oracle/_ref/libfparser_ref.so — the reference generates this file from
// fpconfig.hh.in with its build system, which is not run here.  Only the double parser is built, without the optimizer (OpenRAVE
// never calls Optimize() on a mimic equation, kinbodyjoint.cpp:2243-2249).
#ifndef ORACLE_FPCONFIG_HH
#define ORACLE_FPCONFIG_HH
#define FP_USE_THREAD_SAFE_EVAL
#endif
