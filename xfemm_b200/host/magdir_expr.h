// Evaluator for the magnetisation-direction expressions of block labels (`MagDirFctn`).
// The reference hands the string to its embedded Lua 4.0 interpreter once per element as
//   "x=..\ny=..\nr=x\nz=y\ntheta=..\nR=..\nreturn <expr>"      (static2d.cpp:513-583, staticaxi.cpp:354-420)
// and keeps the real part of the result.  This build has no Lua; it evaluates the expression
// sublanguage those strings are written in — numbers, + - * / ^, unary minus, parentheses, the
// variables x y r z theta R, the constant PI and the math functions of the reference's Lua
// (radians, liblua/lmathlib.cpp:24,280-310) — and rejects anything else with a message.
#pragma once
#include <string>
#include <vector>

namespace xbhost {

struct MagDirVars { double x, y, theta, R; };

class MagDirExpr
{
public:
    // false + message when the string is outside the supported sublanguage
    bool compile(const std::string &text, std::string &error);
    // false when the value is not a finite real number
    bool eval(const MagDirVars &v, double &out) const;
    struct Op { int code; double value; }; // reverse Polish program
private:
    std::vector<Op> ops_;
};

} // namespace xbhost
