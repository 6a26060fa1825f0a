// oracle/ntl_shim/selftest.cpp - TEST INFRASTRUCTURE ONLY.  Prints values of the NTL stand-in at 1000 bits with 60
// digits; tests/test_refapp_cpu.py compares them with mpmath.
#include <NTL/RR.h>
#include <NTL/ZZ.h>
#include <NTL/mat_RR.h>
#include <sstream>
using namespace std;
using namespace NTL;
int main()
{
    RR::SetPrecision(1000);
    RR::SetOutputPrecision(60);
    cout << "pi " << ComputePi_RR() << "\n";
    cout << "cos1 " << cos(RR(1)) << "\n";
    cout << "sin10.5 " << sin(RR(10.5)) << "\n";
    cout << "cos-37.25 " << cos(RR(-37.25)) << "\n";
    cout << "sqrt2 " << sqrt(RR(2)) << "\n";
    cout << "third " << RR(1) / 3 << "\n";
    cout << "e " << exp(RR(1)) << "\n";
    cout << "log10 " << log(RR(10)) << "\n";
    cout << "pow " << pow(RR(2), RR(0.5)) << "\n";
    cout << "3^200 " << power(RR(3), 200) << "\n";
    cout << "rounding " << floor(RR(-2.5)) << " " << ceil(RR(-2.5)) << " " << round(RR(2.5)) << " " << round(RR(3.5)) << " "
         << round(RR(-0.4)) << " " << trunc(RR(-7.9)) << "\n";
    RR x;
    istringstream("  -1.25e-3 ") >> x;
    cout << "parsed " << x << "\n";
    RR::SetOutputPrecision(17);
    cout << "double " << to_double(RR(1) / 3) << " " << to_double(ComputePi_RR()) << "\n";
    mat_RR m, mi;
    m.SetDims(3, 3);
    double v[9] = { 2, 1, 0, 1, 3, 1, 0, 1, 4 };
    for (int i = 0; i < 9; i++)
        m[i / 3][i % 3] = v[i];
    RR d;
    inv(d, mi, m);
    cout << "det " << d << " inv00 " << mi[0][0] << " inv12 " << mi[1][2] << "\n";
    cout << "zz " << (RoundToZZ(RR(7.6)) % 2) << " " << (RoundToZZ(RR(-8.4)) % 2) << "\n";
    RR c = cos(RR(1)), s = sin(RR(1));
    RR::SetOutputPrecision(5);
    cout << "identity " << (abs(c * c + s * s - 1) < power2_RR(-1000) ? "ok" : "bad") << "\n";
    // Hilbert matrix 12x12: inverse times matrix is the identity to ~1000 bits (condition number 1e16)
    mat_RR h, hi;
    h.SetDims(12, 12);
    for (int i = 0; i < 12; i++)
        for (int j = 0; j < 12; j++)
            h[i][j] = RR(1) / (i + j + 1);
    inv(d, hi, h);
    RR worst(0);
    for (int i = 0; i < 12; i++)
        for (int j = 0; j < 12; j++)
        {
            RR acc(0);
            for (int k = 0; k < 12; k++)
                acc += hi[i][k] * h[k][j];
            worst = max(worst, abs(acc - (i == j ? 1 : 0)));
        }
    cout << "hilbert " << (worst < power2_RR(-900) ? "ok" : "bad") << "\n";
    return 0;
}
