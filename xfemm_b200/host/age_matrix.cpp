// Air-gap-element matrix: the 10x10 symmetric matrix of one annular "quad" cell of an AGE ring,
// as a closed form in the mesh shifts ci, co and the aspect ratio K = dr/(R dθ), Ki = 1/K.
// Reference: cfemm/fsolver/static2d.cpp:200-274 (55 polynomials, identical in harmonic2d.cpp).
//
// Only 30 of the 55 entries are independent; this restatement evaluates those and fills the
// rest from the structure of the matrix (verified numerically against all 55 reference
// expressions to 1.4e-15, see tests/test_host_parity.py):
//   nodes 0-4 sit on the inner ring (shift a=ci), nodes 5-9 on the outer ring (shift b=co)
//   inner-inner block   II[i][j] = G_ij(ci)
//   outer-outer block   OO[i][j] = G_ij(co)                (same polynomials, other shift)
//   cross block         IO[i][j](ci,co) = H_ij(ci,co) for i<=j,  = H_ji(co,ci) for i>j
#include <cmath>

namespace xb {

static inline double pw(double x, int n)
{
    double r = 1.0;
    for (int i = 0; i < n; i++) r *= x;
    return r;
}

// upper triangle of the ring-ring block for shift c
static void ring_block(double c, double K, double Ki, double G[5][5])
{
    G[0][0] = (5*pw(-1 + c,2)*pw(c,4)*(K + Ki))/48.;
    G[0][1] = -((-1 + c)*pw(c,3)*(5*(-1 + c*(-5 + 4*c))*K + (-5 + c*(-19 + 14*c))*Ki))/48.;
    G[0][2] = ((-1 + c)*pw(c,2)*(5*(2 + c*(-1 - 9*c + 6*pw(c,2)))*K + (10 + c*(1 + 3*c*(-7 + 4*c)))*Ki))/48.;
    G[0][3] = -(pw(-1 + c,2)*pw(c,2)*(5*(-2 + c*(-3 + 4*c))*K + (2 + c*(-3 + 2*c))*Ki))/48.;
    G[0][4] = (pw(-1 + c,3)*pw(c,3)*(5*K - Ki))/48.;
    G[1][1] = (pw(c,2)*(5*pw(1 + (5 - 4*c)*c,2)*K + (5 + c*(38 + c*(49 + 4*c*(-29 + 11*c))))*Ki))/48.;
    G[1][2] = (-5*c*(-1 + 2*c)*(-2 + 3*(-1 + c)*c)*(-1 + c*(-5 + 4*c))*K + c*(10 + c*(39 - c*(50 + c*(85 + 6*c*(-23 + 8*c)))))*Ki)/48.;
    G[1][3] = ((-1 + c)*c*(5*(2 + c*(13 + c*(3 + 16*(-2 + c)*c)))*K + (-2 + 5*c*(1 + c*(3 + 4*(-2 + c)*c)))*Ki))/48.;
    G[1][4] = -(pw(-1 + c,2)*pw(c,2)*(5*(-1 + c*(-5 + 4*c))*K + Ki + c*(-1 + 2*c)*Ki))/48.;
    G[2][2] = (5*pw(-2 + c + 9*pw(c,2) - 6*pw(c,3),2)*K + (20 + (-1 + c)*c*(-4 + 3*(-1 + c)*c*(-25 + 24*(-1 + c)*c)))*Ki)/48.;
    G[2][3] = (-5*(4 + pw(c,2)*(-33 + c*(18 + c*(65 + 6*c*(-13 + 4*c)))))*K + (4 + pw(c,2)*(39 - c*(30 + c*(115 + 6*c*(-25 + 8*c)))))*Ki)/48.;
    G[2][4] = (pw(-1 + c,2)*c*(5*(2 + c*(-1 - 9*c + 6*pw(c,2)))*K + (-2 + c*(-5 + 3*c*(-5 + 4*c)))*Ki))/48.;
    G[3][3] = (pw(-1 + c,2)*(5*pw(2 + (3 - 4*c)*c,2)*K + (20 + c*(36 + c*(-35 - 60*c + 44*pw(c,2))))*Ki))/48.;
    G[3][4] = -(pw(-1 + c,3)*c*(5*(-2 + c*(-3 + 4*c))*K + (-10 + c*(-9 + 14*c))*Ki))/48.;
    G[4][4] = (5*pw(-1 + c,4)*pw(c,2)*(K + Ki))/48.;
}

// H_ij(a,b) for i<=j: inner node i (shift a) against outer node j (shift b)
static void cross_upper(double a, double b, double K, double Ki, double H[5][5])
{
    H[0][0] = ((-1 + a)*pw(a,2)*(-1 + b)*pw(b,2)*(K - 5*Ki))/48.;
    H[0][1] = -((-1 + a)*pw(a,2)*b*((-1 + b*(-5 + 4*b))*K + (5 + (19 - 14*b)*b)*Ki))/48.;
    H[0][2] = ((-1 + a)*pw(a,2)*((2 + b*(-1 - 9*b + 6*pw(b,2)))*K - (10 + b*(1 + 3*b*(-7 + 4*b)))*Ki))/48.;
    H[0][3] = -((-1 + a)*pw(a,2)*(-1 + b)*((-2 + b*(-3 + 4*b))*K + (-2 + (3 - 2*b)*b)*Ki))/48.;
    H[0][4] = ((-1 + a)*pw(a,2)*pw(-1 + b,2)*b*(K + Ki))/48.;
    H[1][1] = (a*b*((-1 + a*(-5 + 4*a))*(-1 + b*(-5 + 4*b))*K + (-5 + a*(-19 + 14*a) - 19*b + a*(-77 + 58*a)*b + 2*(7 + (29 - 22*a)*a)*pw(b,2))*Ki))/48.;
    H[1][2] = (-(a*(-1 + a*(-5 + 4*a))*(2 + b*(-1 - 9*b + 6*pw(b,2)))*K) + a*(-10 + b*(-1 + 3*(7 - 4*b)*b) + a*(-38 + b + 99*pw(b,2) - 60*pw(b,3)) + pw(a,2)*(28 + 2*b*(-1 + 3*b*(-13 + 8*b))))*Ki)/48.;
    H[1][3] = (a*(-1 + b)*((-1 + a*(-5 + 4*a))*(-2 + b*(-3 + 4*b))*K + (2 + b*(-3 + 2*b) + pw(a,2)*(4 + 2*(9 - 10*b)*b) + a*(-2 + b*(-21 + 22*b)))*Ki))/48.;
    H[1][4] = -(a*pw(-1 + b,2)*b*((-1 + a*(-5 + 4*a))*K + (-1 + a - 2*pw(a,2))*Ki))/48.;
    H[2][2] = ((2 + a*(-1 - 9*a + 6*pw(a,2)))*(2 + b*(-1 - 9*b + 6*pw(b,2)))*K - (2*(10 + b) + 6*pw(b,2)*(-7 + 4*b) + 3*pw(a,2)*(-14 + b*(5 + (55 - 36*b)*b)) + a*(2 + b*(5 + 3*(5 - 4*b)*b)) + 12*pw(a,3)*(2 + b*(-1 - 9*b + 6*pw(b,2))))*Ki)/48.;
    H[2][3] = (-((2 + a*(-1 - 9*a + 6*pw(a,2)))*(2 + b - 7*pw(b,2) + 4*pw(b,3))*K) + (-1 + b)*(4 + 2*a*(5 + 3*(5 - 4*a)*a) + 3*(-2 + a*(3 + (17 - 12*a)*a))*b + 2*(2 + a*(-7 + 3*a*(-11 + 8*a)))*pw(b,2))*Ki)/48.;
    H[2][4] = (pw(-1 + b,2)*b*((2 + a*(-1 - 9*a + 6*pw(a,2)))*K + (2 + a*(5 + 3*(5 - 4*a)*a))*Ki))/48.;
    H[3][3] = ((-1 + a)*(-1 + b)*((-2 + a*(-3 + 4*a))*(-2 + b*(-3 + 4*b))*K + (-20 + 3*a*(1 + 2*b)*(-6 + 5*b) + 2*b*(-9 + 14*b) + pw(a,2)*(28 + 30*b - 44*pw(b,2)))*Ki))/48.;
    H[3][4] = -((-1 + a)*pw(-1 + b,2)*b*((-2 + a*(-3 + 4*a))*K + (10 + (9 - 14*a)*a)*Ki))/48.;
    H[4][4] = (pw(-1 + a,2)*a*pw(-1 + b,2)*b*(K - 5*Ki))/48.;
}

// MG: row-major 10x10, full symmetric fill
void age_cell_matrix(double ci, double co, double K, double *MG)
{
    const double Ki = 1. / K;
    double Gi[5][5], Go[5][5], Hab[5][5], Hba[5][5];
    ring_block(ci, K, Ki, Gi);
    ring_block(co, K, Ki, Go);
    cross_upper(ci, co, K, Ki, Hab);
    cross_upper(co, ci, K, Ki, Hba);
    for (int i = 0; i < 5; i++)
        for (int j = 0; j < 5; j++)
        {
            const double gi = (i <= j) ? Gi[i][j] : Gi[j][i];
            const double go = (i <= j) ? Go[i][j] : Go[j][i];
            const double h = (i <= j) ? Hab[i][j] : Hba[j][i];
            MG[i * 10 + j] = gi;
            MG[(5 + i) * 10 + (5 + j)] = go;
            MG[i * 10 + (5 + j)] = h;
            MG[(5 + j) * 10 + i] = h;
        }
}

} // namespace xb
