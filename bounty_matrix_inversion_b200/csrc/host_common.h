// Small host helpers shared by the client and the engine translation units.
#pragma once
#include <string>
#include <vector>

#include "field.cuh"

namespace bmi_host {

void set_error(const std::string& msg);

inline int brv(int x, int L) {
    int r = 0;
    for (int b = 0; b < L; b++)
        if (x >> b & 1) r |= 1 << (L - 1 - b);
    return r;
}

// tw[i] = psi^brv(i), twi[i] = psi^-brv(i), psi = 7^((p-1)/2N) a primitive 2N-th root of unity
inline void twiddles(int N, std::vector<u64>& tw, std::vector<u64>& twi) {
    int L = 0;
    while ((1 << L) < N) L++;
    const u64 psi = fpow(7, (BMI_P - 1) / (2 * (u64)N)), psii = fpow(psi, BMI_P - 2);
    std::vector<u64> pw(N), pwi(N);
    pw[0] = pwi[0] = 1;
    for (int i = 1; i < N; i++) { pw[i] = fmul(pw[i - 1], psi); pwi[i] = fmul(pwi[i - 1], psii); }
    tw.resize(N);
    twi.resize(N);
    for (int i = 0; i < N; i++) { tw[i] = pw[brv(i, L)]; twi[i] = pwi[brv(i, L)]; }
}

}  // namespace bmi_host
