// Probe (GPU box only): accuracy of sin.approx / cos.approx for arguments of many cycles, i.e. whether the
// MUFU range reduction makes the explicit "p - rint(p)" of the steering synthesis unnecessary.
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__global__ void probe(int n, float pmax, double *err_direct, double *err_reduced) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float p = -pmax + 2.0f * pmax * (float(i) + 0.37f) / float(n);      // phase in cycles
    double sd, cd;
    sincospi(2.0 * double(p), &sd, &cd);
    float a = 6.28318530717958647692f * p, s1, c1;
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s1) : "f"(a));
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c1) : "f"(a));
    float rn = (p + 12582912.0f) - 12582912.0f, fr = p - rn, s2, c2;
    __sincosf(6.28318530717958647692f * fr, &s2, &c2);
    double e1 = fmax(fabs(double(s1) - sd), fabs(double(c1) - cd));
    double e2 = fmax(fabs(double(s2) - sd), fabs(double(c2) - cd));
    // atomic max on doubles via unsigned long long compare (values are non-negative)
    atomicMax((unsigned long long *)err_direct, (unsigned long long)__double_as_longlong(e1));
    atomicMax((unsigned long long *)err_reduced, (unsigned long long)__double_as_longlong(e2));
}
int main() {
    double *d; cudaMalloc(&d, 16);
    for (float pmax : {0.5f, 4.0f, 32.0f, 256.0f, 4096.0f}) {
        cudaMemset(d, 0, 16);
        probe<<<4096, 256>>>(1 << 20, pmax, d, d + 1);
        double h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("|p| <= %7.1f cycles: max abs err direct sin/cos.approx(2 pi p) = %.3e, explicit reduction = %.3e\n", pmax, h[0], h[1]);
    }
    return 0;
}
