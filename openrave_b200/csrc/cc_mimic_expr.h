// Host-side evaluator for mimic equations that the device chain program cannot hold (anything but a * q[dof] + b).
//
// The reference parses a mimic_pos string with the vendored fparser 4.5.2 after replacing joint names by variables
// (kinbodyjoint.cpp:2236-2249, CreateJointFunctionParser :748-765) and evaluates it inside every SetDOFValues
// (kinbody.cpp:2308-2320, Joint::_Eval kinbodyjoint.cpp:2492-2497).  This is a from-scratch recursive-descent compiler for the same
// language into a small stack program: fparser's operator table and precedences ("The function string syntax" of its
// documentation: () ^ unary- ! * / % + - comparisons & |, '^' right-associative and binding tighter than a unary minus on its
// left), its elementary functions, its comparison epsilon as OpenRAVE sets it (g_fEpsilonLinear, kinbodyjoint.cpp:751-753) and its
// truth rule (|x| >= 0.5), plus OpenRAVE's single-valued triangle helpers SSSA / SASA / SASS.  Not supported, and refused at parse
// time: inline variables (y := ...), units, user constants, and OpenRAVE's multi-valued polyroots* helpers.
//
// Pinned in tests/test_mimic_expr.py against the vendored fparser itself (oracle/_ref/libfparser_ref.so).
#pragma once

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "cc_epsilons.h"

namespace ccexpr {

enum Code : int {
    kConst, kVar, kNeg, kNot, kAdd, kSub, kMul, kDiv, kMod, kPow, kEq, kNe, kLt, kLe, kGt, kGe, kAnd, kOr, kJumpIfFalse, kJump,
    kFn1, kFn2, kFn3
};
enum Fn : int {
    fAbs, fAcos, fAcosh, fAsin, fAsinh, fAtan, fAtanh, fCbrt, fCeil, fCos, fCosh, fCot, fCsc, fExp, fExp2, fFloor, fInt, fLog, fLog10,
    fLog2, fSec, fSin, fSinh, fSqrt, fTan, fTanh, fTrunc,                     // one argument
    fAtan2, fHypot, fMax, fMin, fPowFn,                                         // two arguments
    fSSSA, fSASA, fSASS                                                         // three: OpenRAVE's triangle helpers (kinbodyjoint.cpp:722-746)
};
struct FnName { const char* name; int fn; int nargs; };
static const FnName kFunctions[] = {
    {"abs", fAbs, 1}, {"acos", fAcos, 1}, {"acosh", fAcosh, 1}, {"asin", fAsin, 1}, {"asinh", fAsinh, 1}, {"atan", fAtan, 1},
    {"atan2", fAtan2, 2}, {"atanh", fAtanh, 1}, {"cbrt", fCbrt, 1}, {"ceil", fCeil, 1}, {"cos", fCos, 1}, {"cosh", fCosh, 1},
    {"cot", fCot, 1}, {"csc", fCsc, 1}, {"exp", fExp, 1}, {"exp2", fExp2, 1}, {"floor", fFloor, 1}, {"hypot", fHypot, 2},
    {"int", fInt, 1}, {"log", fLog, 1}, {"log10", fLog10, 1}, {"log2", fLog2, 1}, {"max", fMax, 2}, {"min", fMin, 2},
    {"pow", fPowFn, 2}, {"sec", fSec, 1}, {"sin", fSin, 1}, {"sinh", fSinh, 1}, {"sqrt", fSqrt, 1}, {"tan", fTan, 1},
    {"tanh", fTanh, 1}, {"trunc", fTrunc, 1},
    // registered on top of fparser by CreateJointFunctionParser (kinbodyjoint.cpp:748-765); COLLADA formulas of closed linkages
    // use them (colladareader.cpp:5571-5574)
    {"SSSA", fSSSA, 3}, {"SASA", fSASA, 3}, {"SASS", fSASS, 3},
};

// fparser's EvalError() codes, plus kNotFinite: a power or the final value that is NaN or infinite.  (fparser rewrites powers at
// parse time — x^0.5 into sqrt, x^-2 into 1/(x*x), x^y into exp(y*log x) — and which of its codes, or a silent NaN / inf, such a
// case produces depends on the rewrite; here every one of them is an error.  x^0.5 and x^(1/3) are compiled to sqrt / cbrt as
// fparser does, so a negative base is code 2 for the first and fine for the second.)
enum EvalError : int { kOk = 0, kDivisionByZero = 1, kSqrtOfNegative = 2, kLogOfNonPositive = 3, kTrigDomain = 4, kNotFinite = 5 };

struct Op { int code; int arg; double value; };

// x^y the way fparser evaluates it (fp_pow in its fpaux.hh): whole exponents by repeated squaring, positive bases through
// exp(y log x), and — its own convention — a negative base with an exponent that is not a multiple of 1/16 gives -(|x|^y) where
// the C library says NaN.  Everything else is pow().
inline double power(double x, double y)
{
    if (x == 1.0) return 1.0;
    if (fabs(y) < 9e18 && y == (double)(long long)y) {
        unsigned long long n = (unsigned long long)(y < 0 ? -(long long)y : (long long)y);
        double r = 1.0, b = x;
        while (n) { if (n & 1) { r *= b; n -= 1; } else { b *= b; n /= 2; } }
        return y < 0 ? 1.0 / r : r;
    }
    const double y16 = y * 16.0;
    const bool sixteenth = y16 == floor(y16);
    if (y >= 0) {
        if (x > 0) return exp(log(x) * y);
        if (x == 0) return 0.0;
        if (!sixteenth) return -exp(log(-x) * y);
    } else {
        if (x > 0) return exp(log(1.0 / x) * -y);
        if (x < 0 && !sixteenth) return -exp(log(-1.0 / x) * -y);
    }
    return pow(x, y);
}

struct Program {
    std::vector<Op> ops;
    std::vector<int> vars;        // distinct variable ids the expression reads, in order of first use
    int depth = 0;                // stack slots the program needs

    // g_fEpsilonLinear (libopenrave.h:141-142): OpenRAVE hands it to fparser as the comparison epsilon (kinbodyjoint.cpp:751-753)
    static double epsilon() { return CC_EPSILON_LINEAR; }
    static bool truth(double x) { return fabs(x) >= 0.5; }

    // f(values[id] for every variable id); *err receives the first EvalError (the result is then 0, as fparser returns)
    double eval(const double* values, int* err = nullptr) const
    {
        double stack_small[32];
        std::vector<double> stack_big;
        double* st = stack_small;
        if (depth > 32) { stack_big.resize((size_t)depth); st = stack_big.data(); }
        int sp = 0, e = kOk;
        const double eps = epsilon();
        for (size_t pc = 0; pc < ops.size() && e == kOk; ++pc) {
            const Op& o = ops[pc];
            switch (o.code) {
            case kConst: st[sp++] = o.value; break;
            case kVar: st[sp++] = values[o.arg]; break;
            case kNeg: st[sp - 1] = -st[sp - 1]; break;
            case kNot: st[sp - 1] = truth(st[sp - 1]) ? 0.0 : 1.0; break;
            case kJumpIfFalse: if (!truth(st[--sp])) pc = (size_t)o.arg - 1; break;
            case kJump: pc = (size_t)o.arg - 1; break;
            case kFn1: {
                const double x = st[sp - 1];
                double r = 0;
                switch (o.arg) {
                case fAbs: r = fabs(x); break;
                case fAcos: if (x < -1 || x > 1) e = kTrigDomain; else r = acos(x); break;
                case fAcosh: if (x < 1) e = kTrigDomain; else r = acosh(x); break;
                case fAsin: if (x < -1 || x > 1) e = kTrigDomain; else r = asin(x); break;
                case fAsinh: r = asinh(x); break;
                case fAtan: r = atan(x); break;
                case fAtanh: if (x <= -1 || x >= 1) e = kTrigDomain; else r = atanh(x); break;
                case fCbrt: r = cbrt(x); break;
                case fCeil: r = ceil(x); break;
                case fCos: r = cos(x); break;
                case fCosh: r = cosh(x); break;
                case fCot: { const double t = tan(x); if (t == 0) e = kDivisionByZero; else r = 1.0 / t; break; }
                case fCsc: { const double s = sin(x); if (s == 0) e = kDivisionByZero; else r = 1.0 / s; break; }
                case fExp: r = exp(x); break;
                case fExp2: r = power(2.0, x); break;
                case fFloor: r = floor(x); break;
                case fInt: r = x < 0 ? ceil(x - 0.5) : floor(x + 0.5); break;            // to nearest, ties away from zero
                case fLog: if (x <= 0) e = kLogOfNonPositive; else r = log(x); break;
                case fLog10: if (x <= 0) e = kLogOfNonPositive; else r = log10(x); break;
                case fLog2: if (x <= 0) e = kLogOfNonPositive; else r = log2(x); break;
                case fSec: { const double c = cos(x); if (c == 0) e = kDivisionByZero; else r = 1.0 / c; break; }
                case fSin: r = sin(x); break;
                case fSinh: r = sinh(x); break;
                case fSqrt: if (x < 0) e = kSqrtOfNegative; else r = sqrt(x); break;
                case fTan: r = tan(x); break;
                case fTanh: r = tanh(x); break;
                case fTrunc: r = trunc(x); break;
                }
                st[sp - 1] = r;
                break;
            }
            case kFn3: {
                const double c = st[--sp], b = st[--sp], a = st[sp - 1];
                double r = 0;
                switch (o.arg) {
                case fSSSA: {                 // three sides -> the angle opposite the first
                    const double f = (a * a + b * b - c * c) / (2 * b);
                    r = atan2(sqrt(a * a - f * f), b - f);
                    break;
                }
                case fSASA: r = atan2(a * sin(b), c - a * cos(b)); break;            // side, angle, side -> the missing angle
                case fSASS: r = sqrt(a * a + c * c - 2 * a * c * cos(b)); break;      // side, angle, side -> the missing side
                }
                if (!isfinite(r)) { e = kNotFinite; r = 0; }
                st[sp - 1] = r;
                break;
            }
            default: {
                const double b = st[--sp], a = st[sp - 1];
                double r = 0;
                switch (o.code) {
                case kAdd: r = a + b; break;
                case kSub: r = a - b; break;
                case kMul: r = a * b; break;
                case kDiv: if (b == 0) e = kDivisionByZero; else r = a / b; break;
                case kMod: if (b == 0) e = kDivisionByZero; else r = fmod(a, b); break;
                case kPow: r = power(a, b); if (!isfinite(r)) { e = kNotFinite; r = 0; } break;
                case kEq: r = fabs(a - b) <= eps; break;
                case kNe: r = fabs(a - b) > eps; break;
                case kLt: r = a < b - eps; break;
                case kLe: r = a <= b + eps; break;
                case kGt: r = b < a - eps; break;
                case kGe: r = b <= a + eps; break;
                case kAnd: r = truth(a) && truth(b); break;
                case kOr: r = truth(a) || truth(b); break;
                case kFn2:
                    switch (o.arg) {
                    case fAtan2: r = atan2(a, b); break;
                    case fHypot: r = hypot(a, b); break;
                    case fMax: r = a < b ? b : a; break;
                    case fMin: r = b < a ? b : a; break;
                    case fPowFn: r = power(a, b); if (!isfinite(r)) { e = kNotFinite; r = 0; } break;
                    }
                    break;
                }
                st[sp - 1] = r;
                break;
            }
            }
        }
        if (e == kOk && (sp != 1 || !isfinite(st[0]))) e = kNotFinite;
        if (err) *err = e;
        return e == kOk ? st[0] : 0.0;
    }
};

// Compiles `text` over the variable names `names[0..nnames)` (joint names; a name is matched as a whole word, longest first,
// which is what the reference's name substitution amounts to).  On failure returns false with the position and a message.
class Compiler {
public:
    Compiler(const char* text, const char* const* names, int nnames) : s_(text ? text : ""), names_(names), nnames_(nnames) {}

    bool run(Program& out, int* error_pos, std::string* error)
    {
        prog_ = &out;
        out.ops.clear(); out.vars.clear(); out.depth = 0;
        pos_ = 0; cur_ = 0; failed_ = false; barrier_ = 0;
        skip();
        if (!s_[pos_]) fail("empty expression");
        else {
            parse_or();
            skip();
            if (!failed_ && s_[pos_]) fail(s_[pos_] == ':' || s_[pos_] == ';' ? "inline variables are not supported" : "syntax error");
        }
        if (failed_) {
            if (error_pos) *error_pos = (int)errpos_;
            if (error) *error = msg_;
            out.ops.clear();
            return false;
        }
        return true;
    }

private:
    const char* s_;
    const char* const* names_;
    int nnames_;
    Program* prog_ = nullptr;
    size_t pos_ = 0, errpos_ = 0, barrier_ = 0;    // barrier_: index of the first op after the last if()
    int cur_ = 0;                 // stack depth at this point of the emitted program
    bool failed_ = false;
    std::string msg_;

    static bool word(char c) { return (c >= '0' && c <= '9') || (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c == '_'; }
    void skip() { while (s_[pos_] == ' ' || s_[pos_] == '\t' || s_[pos_] == '\n' || s_[pos_] == '\r') ++pos_; }
    void fail(const char* m) { if (!failed_) { failed_ = true; errpos_ = pos_; msg_ = m; } }
    void emit(int code, int arg = 0, double value = 0.0, int pops = 0, int pushes = 0)
    {
        std::vector<Op>& ops = prog_->ops;
        const size_t n = ops.size();
        // constants fold (so that x^(1/2) and x^(1/3) are seen for what they are), and powers with those two exponents become the
        // root functions, as in fparser.  Two constants in front of a binary operator are its operands unless an if() ended between
        // or right after them (its branches join there): nothing is rewritten across that point.
        if (pops == 2 && pushes == 1 && n >= 2 + barrier_ && ops[n - 1].code == kConst && ops[n - 2].code == kConst && code >= kAdd && code <= kPow) {
            Program two;
            two.ops = {ops[n - 2], ops[n - 1], Op{code, arg, value}};
            two.depth = 2;
            int err = 0;
            const double v = two.eval(nullptr, &err);
            if (err == kOk) { ops.pop_back(); ops.back().value = v; cur_ -= 1; return; }
        }
        if ((code == kPow || (code == kFn2 && arg == fPowFn)) && n >= 1 + barrier_ && ops[n - 1].code == kConst && (ops[n - 1].value == 0.5 || ops[n - 1].value == 1.0 / 3.0)) {
            const int fn = ops[n - 1].value == 0.5 ? fSqrt : fCbrt;
            ops.pop_back();
            ops.push_back(Op{kFn1, fn, 0.0});
            cur_ -= 1;
            return;
        }
        ops.push_back(Op{code, arg, value});
        cur_ += pushes - pops;
        if (cur_ > prog_->depth) prog_->depth = cur_;
    }
    bool eat(char c) { skip(); if (s_[pos_] == c) { ++pos_; return true; } return false; }

    void parse_or()
    {
        parse_and();
        for (skip(); !failed_ && s_[pos_] == '|'; skip()) { ++pos_; parse_and(); emit(kOr, 0, 0, 2, 1); }
    }
    void parse_and()
    {
        parse_cmp();
        for (skip(); !failed_ && s_[pos_] == '&'; skip()) { ++pos_; parse_cmp(); emit(kAnd, 0, 0, 2, 1); }
    }
    void parse_cmp()
    {
        parse_add();
        for (skip(); !failed_; skip()) {
            int code;
            if (s_[pos_] == '<' && s_[pos_ + 1] == '=') { code = kLe; pos_ += 2; }
            else if (s_[pos_] == '>' && s_[pos_ + 1] == '=') { code = kGe; pos_ += 2; }
            else if (s_[pos_] == '!' && s_[pos_ + 1] == '=') { code = kNe; pos_ += 2; }
            else if (s_[pos_] == '<') { code = kLt; pos_ += 1; }
            else if (s_[pos_] == '>') { code = kGt; pos_ += 1; }
            else if (s_[pos_] == '=') { code = kEq; pos_ += 1; }
            else break;
            parse_add();
            emit(code, 0, 0, 2, 1);
        }
    }
    void parse_add()
    {
        parse_mul();
        for (skip(); !failed_ && (s_[pos_] == '+' || s_[pos_] == '-'); skip()) {
            const int code = s_[pos_++] == '+' ? kAdd : kSub;
            parse_mul();
            emit(code, 0, 0, 2, 1);
        }
    }
    void parse_mul()
    {
        parse_unary();
        for (skip(); !failed_ && (s_[pos_] == '*' || s_[pos_] == '/' || s_[pos_] == '%'); skip()) {
            const char c = s_[pos_++];
            parse_unary();
            emit(c == '*' ? kMul : c == '/' ? kDiv : kMod, 0, 0, 2, 1);
        }
    }
    void parse_unary()
    {
        skip();
        if (failed_) return;
        if (s_[pos_] == '-') { ++pos_; parse_unary(); emit(kNeg, 0, 0, 1, 1); return; }
        if (s_[pos_] == '!' && s_[pos_ + 1] != '=') { ++pos_; parse_unary(); emit(kNot, 0, 0, 1, 1); return; }
        parse_atom();
        skip();
        if (!failed_ && s_[pos_] == '^') { ++pos_; parse_unary(); emit(kPow, 0, 0, 2, 1); }     // right-associative; 2^-x is fine
    }
    void parse_atom()
    {
        skip();
        if (failed_) return;
        const char c = s_[pos_];
        if (c == '(') {
            ++pos_;
            parse_or();
            if (!failed_ && !eat(')')) fail("expected ')'");
            return;
        }
        if ((c >= '0' && c <= '9') || (c == '.' && s_[pos_ + 1] >= '0' && s_[pos_ + 1] <= '9')) {
            char* end = nullptr;
            const double v = strtod(s_ + pos_, &end);
            if (end == s_ + pos_) { fail("bad number"); return; }
            pos_ = (size_t)(end - s_);
            if (word(s_[pos_])) { fail("syntax error"); return; }                                // "2x": no implicit product
            emit(kConst, 0, v, 0, 1);
            return;
        }
        // a joint name, as a whole word, longest first
        int best = -1; size_t bestlen = 0;
        for (int k = 0; k < nnames_; ++k) {
            const char* nm = names_[k];
            const size_t len = nm ? strlen(nm) : 0;
            if (len > bestlen && strncmp(s_ + pos_, nm, len) == 0 && !(word(nm[len - 1]) && word(s_[pos_ + len]))) { best = k; bestlen = len; }
        }
        if (best >= 0 && !(pos_ > 0 && word(s_[pos_ - 1]) && word(s_[pos_]))) {
            pos_ += bestlen;
            bool seen = false;
            for (int v : prog_->vars) seen = seen || v == best;
            if (!seen) prog_->vars.push_back(best);
            emit(kVar, best, 0, 0, 1);
            return;
        }
        if (!((c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c == '_')) { fail(c ? "syntax error" : "unexpected end of expression"); return; }
        size_t end = pos_;
        while (word(s_[end])) ++end;
        const std::string id(s_ + pos_, end - pos_);
        if (id == "if") { pos_ = end; parse_if(); return; }
        for (const FnName& f : kFunctions)
            if (id == f.name) {
                pos_ = end;
                if (!eat('(')) { fail("expected '(' after a function name"); return; }
                parse_or();
                for (int a = 1; a < f.nargs; ++a) {
                    if (!failed_ && !eat(',')) { fail("expected ','"); return; }
                    parse_or();
                }
                if (!failed_ && !eat(')')) { fail("expected ')'"); return; }
                if (f.nargs == 1) emit(kFn1, f.fn, 0, 1, 1); else if (f.nargs == 2) emit(kFn2, f.fn, 0, 2, 1); else emit(kFn3, f.fn, 0, 3, 1);
                return;
            }
        fail("unknown name (not a joint of this body and not a supported function)");
    }
    void parse_if()
    {
        // if(c, a, b): only the taken branch is evaluated, as in fparser (an error in the other one does not count)
        if (!eat('(')) { fail("expected '(' after if"); return; }
        parse_or();
        if (!failed_ && !eat(',')) { fail("expected ','"); return; }
        if (failed_) return;
        const size_t jz = prog_->ops.size();
        emit(kJumpIfFalse, 0, 0, 1, 0);
        const int before = cur_;
        parse_or();
        if (!failed_ && !eat(',')) { fail("expected ','"); return; }
        if (failed_) return;
        const size_t jmp = prog_->ops.size();
        emit(kJump);
        prog_->ops[jz].arg = (int)prog_->ops.size();
        cur_ = before;                                   // the two branches leave their value in the same slot
        parse_or();
        if (!failed_ && !eat(')')) { fail("expected ')'"); return; }
        if (failed_) return;
        prog_->ops[jmp].arg = (int)prog_->ops.size();
        barrier_ = prog_->ops.size();
    }
};

inline bool compile(const char* text, const char* const* names, int nnames, Program& out, int* error_pos, std::string* error)
{
    return Compiler(text, names, nnames).run(out, error_pos, error);
}

}   // namespace ccexpr
