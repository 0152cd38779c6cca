// Geometry / steering helpers shared by the map kernels.
#pragma once
#include "common.cuh"

// microphones are padded to a multiple of 32 (zero amplitude, zero matrix rows/columns)
__host__ __device__ inline int das_padded_mics(int M) { return (M + 31) / 32 * 32; }

__host__ __device__ inline double das_distance(double ax, double ay, double az, double bx, double by, double bz) {
    const double dx = ax - bx, dy = ay - by, dz = az - bz;
    return sqrt(dx * dx + dy * dy + dz * dz);
}

// |h_m| of the four formulations of beamforming/beamforming.py:1515-1703 (N = number of mics,
// r = |grid - mic_m|, r0 = |grid - reference mic|, s = sum_j 1/r_j^2).
__host__ __device__ inline double das_amplitude(int steering, double r, double r0, double s, int N) {
    switch (steering) {
    case B2_STEER_CLASSIC: return 1.0 / double(N);                    // :1553-1557
    case B2_STEER_INVERSE: return r / (double(N) * r0);               // :1598-1604
    case B2_STEER_TRUE_POWER: return 1.0 / (r0 * r * s);              // :1647-1653
    default: return 1.0 / (r * sqrt(double(N) * s));                  // :1694-1702 (true location)
    }
}

__device__ __forceinline__ void das_sincospi(float x, float *s, float *c) { sincospif(x, s, c); }
__device__ __forceinline__ void das_sincospi(double x, double *s, double *c) { sincospi(x, s, c); }
