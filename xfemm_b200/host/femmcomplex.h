// Minimal complex value type with the public face the callers of the reference rely on
// (members .re/.im, Re()/Im(), arithmetic; reference: cfemm/libfemm/liblua/femmcomplex.h:31).
// Layout-compatible with double2 / std::complex<double> so host arrays go to the device as-is.
#ifndef XFEMM_B200_FEMMCOMPLEX_H
#define XFEMM_B200_FEMMCOMPLEX_H

#include <cmath>
#include <complex>

class CComplex
{
public:
    double re, im;

    CComplex() : re(0), im(0) {}
    CComplex(double x) : re(x), im(0) {}
    CComplex(int x) : re(x), im(0) {}
    CComplex(double x, double y) : re(x), im(y) {}
    CComplex(const std::complex<double> &z) : re(z.real()), im(z.imag()) {}
    operator std::complex<double>() const { return std::complex<double>(re, im); }

    double Re() const { return re; }
    double Im() const { return im; }
    double Abs() const { return std::hypot(re, im); }
    double Arg() const { return std::atan2(im, re); }
    CComplex Conj() const { return CComplex(re, -im); }

    CComplex operator-() const { return CComplex(-re, -im); }
    CComplex &operator+=(const CComplex &z) { re += z.re; im += z.im; return *this; }
    CComplex &operator-=(const CComplex &z) { re -= z.re; im -= z.im; return *this; }
    CComplex &operator*=(const CComplex &z) { double r = re * z.re - im * z.im; im = re * z.im + im * z.re; re = r; return *this; }
    CComplex &operator/=(const CComplex &z) { *this = CComplex(std::complex<double>(*this) / std::complex<double>(z)); return *this; }
    bool operator==(const CComplex &z) const { return re == z.re && im == z.im; }
    bool operator!=(const CComplex &z) const { return !(*this == z); }
};

inline CComplex operator+(CComplex a, const CComplex &b) { return a += b; }
inline CComplex operator-(CComplex a, const CComplex &b) { return a -= b; }
inline CComplex operator*(CComplex a, const CComplex &b) { return a *= b; }
inline CComplex operator/(CComplex a, const CComplex &b) { return a /= b; }
inline double Re(const CComplex &z) { return z.re; }
inline double Im(const CComplex &z) { return z.im; }
inline double Re(double x) { return x; }
inline double abs(const CComplex &z) { return z.Abs(); }
inline double arg(const CComplex &z) { return z.Arg(); }
inline CComplex conj(const CComplex &z) { return z.Conj(); }
inline CComplex exp(const CComplex &z) { return CComplex(std::exp(std::complex<double>(z))); }
inline CComplex sqrt(const CComplex &z) { return CComplex(std::sqrt(std::complex<double>(z))); }
inline CComplex tanh(const CComplex &z) { return CComplex(std::tanh(std::complex<double>(z))); }

#endif
