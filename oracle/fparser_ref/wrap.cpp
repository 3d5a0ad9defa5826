// TEST INFRASTRUCTURE.  C entry points around the reference's vendored FunctionParser (fparser.hh), the class behind
// OpenRAVEFunctionParserReal that evaluates mimic equations (kinbodyjoint.cpp:748-765, 2243-2249; kinbody.cpp:2308-2320).
// Golden checker for the library's own expression evaluator (openrave_b200/csrc/cc_mimic_expr.h): tests/test_mimic_expr.py.
#include "fparser.hh"

#include <math.h>
#include <string.h>

extern "C" {
// Parse `expr` over the comma-separated variable list `vars`; returns -1 on success, else the position of the error (fparser's
// own convention).  *handle receives the parser.
int fpref_parse(const char* expr, const char* vars, void** handle)
{
    // OpenRAVE's build defines OPENRAVE_FPARSER_SETEPSILON for this fparser version (CMakeLists.txt:607-621) and sets the comparison
    // epsilon to g_fEpsilonLinear = RavePow(g_fEpsilon, 0.9) (kinbodyjoint.cpp:751-753, libopenrave.h:140-142)
    FunctionParser::setEpsilon(pow(1e-15, 0.9));
    FunctionParser* fp = new FunctionParser();
    const int rc = fp->Parse(expr, vars);
    if (rc >= 0) { delete fp; *handle = 0; return rc; }
    *handle = fp;
    return -1;
}
// out[i] = f(values[i][0..nvars)); err[i] = FunctionParser::EvalError() (0 = fine).
void fpref_eval(void* handle, const double* values, int nvars, int n, double* out, int* err)
{
    FunctionParser* fp = (FunctionParser*)handle;
    for (int i = 0; i < n; ++i) {
        out[i] = fp->Eval(values + (size_t)i * nvars);
        if (err) err[i] = fp->EvalError();
    }
}
void fpref_free(void* handle) { delete (FunctionParser*)handle; }
}
