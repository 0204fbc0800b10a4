"""Air-gap-element cell matrix: the 55 closed-form entries MG[i][j](ci, co, K, Ki), i <= j, of
cfemm/fsolver/static2d.cpp:220-274 restated in Python (Power(x, n) -> x**n).  TEST INFRASTRUCTURE ONLY;
the product evaluates a reduced 30-formula form (xfemm_b200/host/age_matrix.cpp) that tests/test_oracle.py
checks against this table."""
import numpy as np


def age_matrix(ci, co, K):
    Ki = 1. / K
    MG = np.zeros((10, 10))
    MG[0, 0] = (5*pw(-1 + ci,2)*pw(ci,4)*(K + Ki))/48.
    MG[0, 1] = -((-1 + ci)*pw(ci,3)*(5*(-1 + ci*(-5 + 4*ci))*K + (-5 + ci*(-19 + 14*ci))*Ki))/48.
    MG[0, 2] = ((-1 + ci)*pw(ci,2)*(5*(2 + ci*(-1 - 9*ci + 6*pw(ci,2)))*K + (10 + ci*(1 + 3*ci*(-7 + 4*ci)))*Ki))/48.
    MG[0, 3] = -(pw(-1 + ci,2)*pw(ci,2)*(5*(-2 + ci*(-3 + 4*ci))*K + (2 + ci*(-3 + 2*ci))*Ki))/48.
    MG[0, 4] = (pw(-1 + ci,3)*pw(ci,3)*(5*K - Ki))/48.
    MG[0, 5] = ((-1 + ci)*pw(ci,2)*(-1 + co)*pw(co,2)*(K - 5*Ki))/48.
    MG[0, 6] = -((-1 + ci)*pw(ci,2)*co*((-1 + co*(-5 + 4*co))*K + (5 + (19 - 14*co)*co)*Ki))/48.
    MG[0, 7] = ((-1 + ci)*pw(ci,2)*((2 + co*(-1 - 9*co + 6*pw(co,2)))*K - (10 + co*(1 + 3*co*(-7 + 4*co)))*Ki))/48.
    MG[0, 8] = -((-1 + ci)*pw(ci,2)*(-1 + co)*((-2 + co*(-3 + 4*co))*K + (-2 + (3 - 2*co)*co)*Ki))/48.
    MG[0, 9] = ((-1 + ci)*pw(ci,2)*pw(-1 + co,2)*co*(K + Ki))/48.
    MG[1, 1] = (pw(ci,2)*(5*pw(1 + (5 - 4*ci)*ci,2)*K + (5 + ci*(38 + ci*(49 + 4*ci*(-29 + 11*ci))))*Ki))/48.
    MG[1, 2] = (-5*ci*(-1 + 2*ci)*(-2 + 3*(-1 + ci)*ci)*(-1 + ci*(-5 + 4*ci))*K + ci*(10 + ci*(39 - ci*(50 + ci*(85 + 6*ci*(-23 + 8*ci)))))*Ki)/48.
    MG[1, 3] = ((-1 + ci)*ci*(5*(2 + ci*(13 + ci*(3 + 16*(-2 + ci)*ci)))*K + (-2 + 5*ci*(1 + ci*(3 + 4*(-2 + ci)*ci)))*Ki))/48.
    MG[1, 4] = -(pw(-1 + ci,2)*pw(ci,2)*(5*(-1 + ci*(-5 + 4*ci))*K + Ki + ci*(-1 + 2*ci)*Ki))/48.
    MG[1, 5] = -(ci*(-1 + co)*pw(co,2)*((-1 + ci*(-5 + 4*ci))*K + (5 + (19 - 14*ci)*ci)*Ki))/48.
    MG[1, 6] = (ci*co*((-1 + ci*(-5 + 4*ci))*(-1 + co*(-5 + 4*co))*K + (-5 + ci*(-19 + 14*ci) - 19*co + ci*(-77 + 58*ci)*co + 2*(7 + (29 - 22*ci)*ci)*pw(co,2))*Ki))/48.
    MG[1, 7] = (-(ci*(-1 + ci*(-5 + 4*ci))*(2 + co*(-1 - 9*co + 6*pw(co,2)))*K) + ci*(-10 + co*(-1 + 3*(7 - 4*co)*co) + ci*(-38 + co + 99*pw(co,2) - 60*pw(co,3)) + pw(ci,2)*(28 + 2*co*(-1 + 3*co*(-13 + 8*co))))*Ki)/48.
    MG[1, 8] = (ci*(-1 + co)*((-1 + ci*(-5 + 4*ci))*(-2 + co*(-3 + 4*co))*K + (2 + co*(-3 + 2*co) + pw(ci,2)*(4 + 2*(9 - 10*co)*co) + ci*(-2 + co*(-21 + 22*co)))*Ki))/48.
    MG[1, 9] = -(ci*pw(-1 + co,2)*co*((-1 + ci*(-5 + 4*ci))*K + (-1 + ci - 2*pw(ci,2))*Ki))/48.
    MG[2, 2] = (5*pw(-2 + ci + 9*pw(ci,2) - 6*pw(ci,3),2)*K + (20 + (-1 + ci)*ci*(-4 + 3*(-1 + ci)*ci*(-25 + 24*(-1 + ci)*ci)))*Ki)/48.
    MG[2, 3] = (-5*(4 + pw(ci,2)*(-33 + ci*(18 + ci*(65 + 6*ci*(-13 + 4*ci)))))*K + (4 + pw(ci,2)*(39 - ci*(30 + ci*(115 + 6*ci*(-25 + 8*ci)))))*Ki)/48.
    MG[2, 4] = (pw(-1 + ci,2)*ci*(5*(2 + ci*(-1 - 9*ci + 6*pw(ci,2)))*K + (-2 + ci*(-5 + 3*ci*(-5 + 4*ci)))*Ki))/48.
    MG[2, 5] = ((-1 + co)*pw(co,2)*((2 + ci*(-1 - 9*ci + 6*pw(ci,2)))*K - (10 + ci*(1 + 3*ci*(-7 + 4*ci)))*Ki))/48.
    MG[2, 6] = (-((2 + ci*(-1 - 9*ci + 6*pw(ci,2)))*co*(-1 + co*(-5 + 4*co))*K) + co*(-10 - 38*co + 28*pw(co,2) + pw(ci,2)*(21 + 99*co - 78*pw(co,2)) + ci*(-1 + co - 2*pw(co,2)) + 12*pw(ci,3)*(-1 + co*(-5 + 4*co)))*Ki)/48.
    MG[2, 7] = ((2 + ci*(-1 - 9*ci + 6*pw(ci,2)))*(2 + co*(-1 - 9*co + 6*pw(co,2)))*K - (2*(10 + co) + 6*pw(co,2)*(-7 + 4*co) + 3*pw(ci,2)*(-14 + co*(5 + (55 - 36*co)*co)) + ci*(2 + co*(5 + 3*(5 - 4*co)*co)) + 12*pw(ci,3)*(2 + co*(-1 - 9*co + 6*pw(co,2))))*Ki)/48.
    MG[2, 8] = (-((2 + ci*(-1 - 9*ci + 6*pw(ci,2)))*(2 + co - 7*pw(co,2) + 4*pw(co,3))*K) + (-1 + co)*(4 + 2*ci*(5 + 3*(5 - 4*ci)*ci) + 3*(-2 + ci*(3 + (17 - 12*ci)*ci))*co + 2*(2 + ci*(-7 + 3*ci*(-11 + 8*ci)))*pw(co,2))*Ki)/48.
    MG[2, 9] = (pw(-1 + co,2)*co*((2 + ci*(-1 - 9*ci + 6*pw(ci,2)))*K + (2 + ci*(5 + 3*(5 - 4*ci)*ci))*Ki))/48.
    MG[3, 3] = (pw(-1 + ci,2)*(5*pw(2 + (3 - 4*ci)*ci,2)*K + (20 + ci*(36 + ci*(-35 - 60*ci + 44*pw(ci,2))))*Ki))/48.
    MG[3, 4] = -(pw(-1 + ci,3)*ci*(5*(-2 + ci*(-3 + 4*ci))*K + (-10 + ci*(-9 + 14*ci))*Ki))/48.
    MG[3, 5] = -((-1 + ci)*(-1 + co)*pw(co,2)*((-2 + ci*(-3 + 4*ci))*K + (-2 + (3 - 2*ci)*ci)*Ki))/48.
    MG[3, 6] = ((-1 + ci)*co*((-2 + ci*(-3 + 4*ci))*(-1 + co*(-5 + 4*co))*K + (2 + ci*(-3 + 2*ci) - 2*co + ci*(-21 + 22*ci)*co + 2*(2 + (9 - 10*ci)*ci)*pw(co,2))*Ki))/48.
    MG[3, 7] = (-((2 + ci - 7*pw(ci,2) + 4*pw(ci,3))*(2 + co*(-1 - 9*co + 6*pw(co,2)))*K) + (-1 + ci)*(4 + 2*co*(5 + 3*(5 - 4*co)*co) + ci*(-6 + 3*co*(3 + (17 - 12*co)*co)) + 2*pw(ci,2)*(2 + co*(-7 + 3*co*(-11 + 8*co))))*Ki)/48.
    MG[3, 8] = ((-1 + ci)*(-1 + co)*((-2 + ci*(-3 + 4*ci))*(-2 + co*(-3 + 4*co))*K + (-20 + 3*ci*(1 + 2*co)*(-6 + 5*co) + 2*co*(-9 + 14*co) + pw(ci,2)*(28 + 30*co - 44*pw(co,2)))*Ki))/48.
    MG[3, 9] = -((-1 + ci)*pw(-1 + co,2)*co*((-2 + ci*(-3 + 4*ci))*K + (10 + (9 - 14*ci)*ci)*Ki))/48.
    MG[4, 4] = (5*pw(-1 + ci,4)*pw(ci,2)*(K + Ki))/48.
    MG[4, 5] = (pw(-1 + ci,2)*ci*(-1 + co)*pw(co,2)*(K + Ki))/48.
    MG[4, 6] = -(pw(-1 + ci,2)*ci*co*((-1 + co*(-5 + 4*co))*K + (-1 + co - 2*pw(co,2))*Ki))/48.
    MG[4, 7] = (pw(-1 + ci,2)*ci*((2 + co*(-1 - 9*co + 6*pw(co,2)))*K + (2 + co*(5 + 3*(5 - 4*co)*co))*Ki))/48.
    MG[4, 8] = -(pw(-1 + ci,2)*ci*(-1 + co)*((-2 + co*(-3 + 4*co))*K + (10 + (9 - 14*co)*co)*Ki))/48.
    MG[4, 9] = (pw(-1 + ci,2)*ci*pw(-1 + co,2)*co*(K - 5*Ki))/48.
    MG[5, 5] = (5*pw(-1 + co,2)*pw(co,4)*(K + Ki))/48.
    MG[5, 6] = -((-1 + co)*pw(co,3)*(5*(-1 + co*(-5 + 4*co))*K + (-5 + co*(-19 + 14*co))*Ki))/48.
    MG[5, 7] = ((-1 + co)*pw(co,2)*(5*(2 + co*(-1 - 9*co + 6*pw(co,2)))*K + (10 + co*(1 + 3*co*(-7 + 4*co)))*Ki))/48.
    MG[5, 8] = -(pw(-1 + co,2)*pw(co,2)*(5*(-2 + co*(-3 + 4*co))*K + (2 + co*(-3 + 2*co))*Ki))/48.
    MG[5, 9] = (pw(-1 + co,3)*pw(co,3)*(5*K - Ki))/48.
    MG[6, 6] = (pw(co,2)*(5*pw(1 + (5 - 4*co)*co,2)*K + (5 + co*(38 + co*(49 + 4*co*(-29 + 11*co))))*Ki))/48.
    MG[6, 7] = (-5*co*(-1 + 2*co)*(-2 + 3*(-1 + co)*co)*(-1 + co*(-5 + 4*co))*K + co*(10 + co*(39 - co*(50 + co*(85 + 6*co*(-23 + 8*co)))))*Ki)/48.
    MG[6, 8] = ((-1 + co)*co*(5*(2 + co*(13 + co*(3 + 16*(-2 + co)*co)))*K + (-2 + 5*co*(1 + co*(3 + 4*(-2 + co)*co)))*Ki))/48.
    MG[6, 9] = -(pw(-1 + co,2)*pw(co,2)*(5*(-1 + co*(-5 + 4*co))*K + Ki + co*(-1 + 2*co)*Ki))/48.
    MG[7, 7] = (5*pw(-2 + co + 9*pw(co,2) - 6*pw(co,3),2)*K + (20 + (-1 + co)*co*(-4 + 3*(-1 + co)*co*(-25 + 24*(-1 + co)*co)))*Ki)/48.
    MG[7, 8] = (-5*(4 + pw(co,2)*(-33 + co*(18 + co*(65 + 6*co*(-13 + 4*co)))))*K + (4 + pw(co,2)*(39 - co*(30 + co*(115 + 6*co*(-25 + 8*co)))))*Ki)/48.
    MG[7, 9] = (pw(-1 + co,2)*co*(5*(2 + co*(-1 - 9*co + 6*pw(co,2)))*K + (-2 + co*(-5 + 3*co*(-5 + 4*co)))*Ki))/48.
    MG[8, 8] = (pw(-1 + co,2)*(5*pw(2 + (3 - 4*co)*co,2)*K + (20 + co*(36 + co*(-35 - 60*co + 44*pw(co,2))))*Ki))/48.
    MG[8, 9] = -(pw(-1 + co,3)*co*(5*(-2 + co*(-3 + 4*co))*K + (-10 + co*(-9 + 14*co))*Ki))/48.
    MG[9, 9] = (5*pw(-1 + co,4)*pw(co,2)*(K + Ki))/48.
    return MG + np.triu(MG, 1).T


def pw(x, n):
    return x ** n
