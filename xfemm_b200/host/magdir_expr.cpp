// See magdir_expr.h.  Recursive-descent parser with Lua's precedences:
//   + -  <  * /  <  unary -  <  ^ (right associative), so that -x^2 = -(x^2) and 2^-1 parses.
#include "magdir_expr.h"
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace xbhost {
namespace {

enum Code
{
    NUM, VX, VY, VTHETA, VR, ADD, SUB, MUL, DIV, POW, NEG,
    F_ABS, F_SIN, F_COS, F_TAN, F_ASIN, F_ACOS, F_ATAN, F_ATAN2, F_CEIL, F_FLOOR, F_MOD, F_SQRT, F_MIN, F_MAX, F_LOG,
    F_LOG10, F_EXP, F_DEG, F_RAD, F_TANH, F_COSH, F_SINH, F_RE, F_IM
};
struct Fn { const char *name; int code; int min_args, max_args; };
const Fn FUNCS[] = {
    {"abs", F_ABS, 1, 1}, {"sin", F_SIN, 1, 1}, {"cos", F_COS, 1, 1}, {"tan", F_TAN, 1, 1}, {"asin", F_ASIN, 1, 1},
    {"acos", F_ACOS, 1, 1}, {"atan", F_ATAN, 1, 1}, {"atan2", F_ATAN2, 2, 2}, {"ceil", F_CEIL, 1, 1}, {"floor", F_FLOOR, 1, 1},
    {"mod", F_MOD, 2, 2}, {"sqrt", F_SQRT, 1, 1}, {"min", F_MIN, 1, 8}, {"max", F_MAX, 1, 8}, {"log", F_LOG, 1, 1},
    {"log10", F_LOG10, 1, 1}, {"exp", F_EXP, 1, 1}, {"deg", F_DEG, 1, 1}, {"rad", F_RAD, 1, 1}, {"tanh", F_TANH, 1, 1},
    {"cosh", F_COSH, 1, 1}, {"sinh", F_SINH, 1, 1}, {"re", F_RE, 1, 1}, {"im", F_IM, 1, 1},
};

struct Parser
{
    const char *s;
    std::vector<MagDirExpr::Op> &out;
    std::string err;
    Parser(const char *text, std::vector<MagDirExpr::Op> &o) : s(text), out(o) {}
    void skip() { while (*s && isspace((unsigned char)*s)) s++; }
    bool fail(const std::string &m) { if (err.empty()) err = m; return false; }
    void emit(int c, double v = 0) { out.push_back({c, v}); }

    bool expr()
    {
        if (!term()) return false;
        for (skip(); *s == '+' || *s == '-'; skip())
        {
            const char op = *s++;
            if (!term()) return false;
            emit(op == '+' ? ADD : SUB);
        }
        return true;
    }
    bool term()
    {
        if (!unary()) return false;
        for (skip(); *s == '*' || *s == '/'; skip())
        {
            const char op = *s++;
            if (!unary()) return false;
            emit(op == '*' ? MUL : DIV);
        }
        return true;
    }
    bool unary()
    {
        skip();
        if (*s == '-') { s++; if (!unary()) return false; emit(NEG); return true; }
        return power();
    }
    bool power()
    {
        if (!primary()) return false;
        skip();
        if (*s == '^')
        {
            s++;
            skip();
            // right operand: a unary minus binds to it (2^-x), then ^ again from the right
            if (*s == '-') { s++; if (!power()) return false; emit(NEG); }
            else if (!power()) return false;
            emit(POW);
        }
        return true;
    }
    bool primary()
    {
        skip();
        if (*s == '(')
        {
            s++;
            if (!expr()) return false;
            skip();
            if (*s != ')') return fail("missing ')'");
            s++;
            return true;
        }
        if (isdigit((unsigned char)*s) || (*s == '.' && isdigit((unsigned char)s[1])))
        {
            char *end = nullptr;
            const double v = strtod(s, &end);
            if (end == s) return fail("bad number");
            s = end;
            emit(NUM, v);
            return true;
        }
        if (isalpha((unsigned char)*s) || *s == '_')
        {
            const char *b = s;
            while (isalnum((unsigned char)*s) || *s == '_') s++;
            const std::string id(b, s);
            skip();
            if (*s == '(')
            {
                s++;
                const Fn *fn = nullptr;
                for (const Fn &f : FUNCS) if (id == f.name) fn = &f;
                if (!fn) return fail("function '" + id + "' is not available");
                int nargs = 0;
                skip();
                if (*s != ')')
                    for (;;)
                    {
                        if (!expr()) return false;
                        nargs++;
                        skip();
                        if (*s == ',') { s++; continue; }
                        break;
                    }
                if (*s != ')') return fail("missing ')' after the arguments of " + id);
                s++;
                if (nargs < fn->min_args || nargs > fn->max_args) return fail("wrong number of arguments for " + id);
                emit(fn->code, nargs);
                return true;
            }
            if (id == "x" || id == "r") emit(VX);
            else if (id == "y" || id == "z") emit(VY);
            else if (id == "theta") emit(VTHETA);
            else if (id == "R") emit(VR);
            else if (id == "PI") emit(NUM, 3.141592653589793238462643383);
            else return fail("unknown name '" + id + "'");
            return true;
        }
        return fail(*s ? std::string("unexpected '") + *s + "'" : std::string("unexpected end of the expression"));
    }
};

} // namespace

bool MagDirExpr::compile(const std::string &text, std::string &error)
{
    ops_.clear();
    Parser p(text.c_str(), ops_);
    bool ok = p.expr();
    if (ok) { p.skip(); if (*p.s) ok = p.fail(std::string("unexpected '") + *p.s + "'"); }
    if (!ok)
    {
        error = "magnetization direction function \"" + text + "\": " + p.err +
                " (supported: numbers, + - * / ^, parentheses, x y r z theta R PI and the math functions)";
        ops_.clear();
    }
    return ok;
}

bool MagDirExpr::eval(const MagDirVars &v, double &out) const
{
    double st[64];
    int sp = 0;
    for (const Op &o : ops_)
    {
        if (sp >= 60) return false;
        switch (o.code)
        {
        case NUM: st[sp++] = o.value; break;
        case VX: st[sp++] = v.x; break;
        case VY: st[sp++] = v.y; break;
        case VTHETA: st[sp++] = v.theta; break;
        case VR: st[sp++] = v.R; break;
        case ADD: sp--; st[sp - 1] += st[sp]; break;
        case SUB: sp--; st[sp - 1] -= st[sp]; break;
        case MUL: sp--; st[sp - 1] *= st[sp]; break;
        case DIV: sp--; st[sp - 1] /= st[sp]; break;
        case POW: sp--; st[sp - 1] = pow(st[sp - 1], st[sp]); break;
        case NEG: st[sp - 1] = -st[sp - 1]; break;
        case F_ABS: st[sp - 1] = fabs(st[sp - 1]); break;
        case F_SIN: st[sp - 1] = sin(st[sp - 1]); break;
        case F_COS: st[sp - 1] = cos(st[sp - 1]); break;
        case F_TAN: st[sp - 1] = tan(st[sp - 1]); break;
        case F_ASIN: st[sp - 1] = asin(st[sp - 1]); break;
        case F_ACOS: st[sp - 1] = acos(st[sp - 1]); break;
        case F_ATAN: st[sp - 1] = atan(st[sp - 1]); break;
        case F_ATAN2: sp--; st[sp - 1] = atan2(st[sp - 1], st[sp]); break;
        case F_CEIL: st[sp - 1] = ceil(st[sp - 1]); break;
        case F_FLOOR: st[sp - 1] = floor(st[sp - 1]); break;
        case F_MOD: sp--; st[sp - 1] = fmod(st[sp - 1], st[sp]); break;
        case F_SQRT: st[sp - 1] = sqrt(st[sp - 1]); break;
        case F_MIN: { int n = (int)o.value; double m = st[sp - n]; for (int k = 1; k < n; k++) m = st[sp - n + k] < m ? st[sp - n + k] : m; sp -= n - 1; st[sp - 1] = m; break; }
        case F_MAX: { int n = (int)o.value; double m = st[sp - n]; for (int k = 1; k < n; k++) m = st[sp - n + k] > m ? st[sp - n + k] : m; sp -= n - 1; st[sp - 1] = m; break; }
        case F_LOG: st[sp - 1] = log(st[sp - 1]); break;
        case F_LOG10: st[sp - 1] = log10(st[sp - 1]); break;
        case F_EXP: st[sp - 1] = exp(st[sp - 1]); break;
        case F_DEG: st[sp - 1] = st[sp - 1] * 180. / 3.141592653589793238462643383; break;
        case F_RAD: st[sp - 1] = st[sp - 1] * 3.141592653589793238462643383 / 180.; break;
        case F_TANH: st[sp - 1] = tanh(st[sp - 1]); break;
        case F_COSH: st[sp - 1] = cosh(st[sp - 1]); break;
        case F_SINH: st[sp - 1] = sinh(st[sp - 1]); break;
        case F_RE: break;
        case F_IM: st[sp - 1] = 0.0; break;
        default: return false;
        }
    }
    if (sp != 1 || !std::isfinite(st[0])) return false;
    out = st[0];
    return true;
}

} // namespace xbhost
