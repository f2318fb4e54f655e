// Device twins of the user callbacks on the reference's hot path
// (dynamics / lqr / erf / is_feasible; contract: reference lqrrt/planner.py:35-59,
// constraints.py:29).  Everything here is register-resident, per-thread code:
// n <= 12 states, m <= 6 controls, no tensor cores.  T is float or double.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "../../include/lqrrt_b200.h"

namespace lq {

template <typename T> struct Num;
template <> struct Num<float> {
    static __host__ __device__ __forceinline__ float inf() { return INFINITY; }
    static __device__ __forceinline__ float two_pi() { return 6.283185307179586f; }
    static __device__ __forceinline__ float inv_two_pi() { return 0.15915494309189535f; }
};
template <> struct Num<double> {
    static __host__ __device__ __forceinline__ double inf() { return (double)INFINITY; }
    static __device__ __forceinline__ double two_pi() { return 6.283185307179586; }
    static __device__ __forceinline__ double inv_two_pi() { return 0.15915494309189535; }
};

// float sine / cosine for the rollout's dependent chain: one Cody-Waite reduction by pi/2 and the two degree-7/8
// minimax polynomials of the usual single-precision kernels (max error ~1 ulp on the reduced range), WITHOUT libm's
// Payne-Hanek path for |a| > 1e5 rad inlined three times per rollout step (it is never taken -- headings are a few
// turns at most -- but costs ~700 instructions of code per step and the instruction-cache misses that go with
// them).  Arguments beyond the reduction's exact range are first folded by whole turns.  double uses libm.
__device__ __forceinline__ void sincos_t(float a, float* s, float* c) {
    if (fabsf(a) > 50000.f) a = fmaf(-6.283185307179586f, rintf(a * 0.15915494309189535f), a);
    const float q = rintf(a * 0.63661977236758138f);            // nearest multiple of pi/2
    const int iq = (int)q;
    float r = fmaf(q, -1.5707962512969970703f, a);              // pi/2 split in three, each product exact inside the FMA
    r = fmaf(q, -7.5497894158615963534e-08f, r);
    r = fmaf(q, -5.3903029534742383927e-15f, r);
    const float r2 = r * r;
    float sp = fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f);
    sp = fmaf(sp, r2, -1.6666654611e-1f);
    sp = fmaf(sp * r2, r, r);
    float cp = fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
    cp = fmaf(cp, r2, 4.166664568298827e-2f);
    cp = fmaf(cp, r2, -0.5f);
    cp = fmaf(cp, r2, 1.0f);
    const float ss = (iq & 1) ? cp : sp, cc = (iq & 1) ? sp : cp;
    *s = (iq & 2) ? -ss : ss;
    *c = ((iq + 1) & 2) ? -cc : cc;
}
__device__ __forceinline__ void sincos_t(double a, double* s, double* c) { sincos(a, s, c); }
__device__ __forceinline__ float atan2_t(float y, float x) { return atan2f(y, x); }
// atan2 for the float32 rollout chain: |error| < 2e-7 rad (odd degree-17 polynomial, Abramowitz & Stegun 4.4.49, on
// min/max in [0, 1] with the SFU reciprocal), no special-case branches.  atan2(0, 0) = 0 regardless of the zeros'
// signs and x = -0 counts as +0 -- callers that need IEEE's signed-zero cases (the reference's v = 0 quirk in the
// rudder term) call atan2_t.
__device__ __forceinline__ float atan2_fast(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    float t = __fdividef(mn, mx);
    t = mx == 0.f ? 0.f : t;
    const float t2 = t * t;
    float p = 0.0028662257f;
    p = fmaf(p, t2, -0.0161657367f);
    p = fmaf(p, t2, 0.0429096138f);
    p = fmaf(p, t2, -0.0752896400f);
    p = fmaf(p, t2, 0.1065626393f);
    p = fmaf(p, t2, -0.1420889944f);
    p = fmaf(p, t2, 0.1999355085f);
    p = fmaf(p, t2, -0.3333314528f);
    float r = fmaf(p * t2, t, t);
    r = ay > ax ? 1.57079637f - r : r;
    r = x < 0.f ? 3.14159274f - r : r;
    return copysignf(r, y);
}
__device__ __forceinline__ double atan2_fast(double y, double x) { return atan2(y, x); }
__device__ __forceinline__ double atan2_t(double y, double x) { return atan2(y, x); }
__device__ __forceinline__ float rint_t(float a) { return rintf(a); }
__device__ __forceinline__ double rint_t(double a) { return rint(a); }
__device__ __forceinline__ float abs_t(float a) { return fabsf(a); }
__device__ __forceinline__ double abs_t(double a) { return fabs(a); }
__device__ __forceinline__ float min_t(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ double min_t(double a, double b) { return fmin(a, b); }
__device__ __forceinline__ float max_t(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double max_t(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ float sqrt_t(float a) { return sqrtf(a); }
__device__ __forceinline__ double sqrt_t(double a) { return sqrt(a); }

// Signed angle a_to - a_from on the circle in the reference's form
// atan2(sg*c - cg*s, cg*c + sg*s) (demos/demo_boat_novice.py:112-117).
template <typename T>
__device__ __forceinline__ T wrap_sc(T s_from, T c_from, T s_to, T c_to) {
    return atan2_fast(s_to * c_from - c_to * s_from, c_to * c_from + s_to * s_from);
}

// ---------------------------------------------------------------------------
// dynamics (reference a5).  step(): x (in) -> xn (out); u is a private copy
// the model may saturate in place, exactly as the demos mutate their copy
// (planner.py:390 passes np.copy(u)), so the caller's recorded u stays raw.
// ---------------------------------------------------------------------------

template <typename T> struct BoatSys {            // demos/demo_boat_{novice,intermediate,advanced}.py
    static constexpr int N = 6, M = 3;
    static constexpr int KIND = LQ_SYS_BOAT;
    T invM[3], Dpos[3], Dneg[3], umax[3], B[12], invB[12], tmax[4];
    T rudder, vpos0, vneg0, inv_vpos0, inv_vneg0;
    int thruster_clip, carlike, use_rudder;

    __host__ void unpack(const double* p) {
        int k = 0;
        for (int i = 0; i < 3; i++) invM[i] = (T)p[k++];
        for (int i = 0; i < 3; i++) Dpos[i] = (T)p[k++];
        for (int i = 0; i < 3; i++) Dneg[i] = (T)p[k++];
        for (int i = 0; i < 3; i++) umax[i] = (T)p[k++];
        for (int i = 0; i < 12; i++) B[i] = (T)p[k++];
        for (int i = 0; i < 12; i++) invB[i] = (T)p[k++];
        for (int i = 0; i < 4; i++) tmax[i] = (T)p[k++];
        rudder = (T)p[k++]; vpos0 = (T)p[k++]; vneg0 = (T)p[k++];
        inv_vpos0 = (T)(1.0 / p[k - 2]); inv_vneg0 = (T)(1.0 / p[k - 1]);
        thruster_clip = p[k++] != 0.0; carlike = p[k++] != 0.0; use_rudder = p[k++] != 0.0;
    }
    static constexpr int NPARAMS = 46;
    static constexpr int HEADING = 2;     // state dim whose sin/cos the rollout caches

    __device__ __forceinline__ void step(const T* x, T* u, T dt, T* xn) const {
        T s, c;
        sincos_t(x[2], &s, &c);
        step_sc(x, u, dt, xn, s, c);
    }
    // (s, c) = sincos(x[HEADING]) supplied by the caller (one evaluation per state in the rollout)
    __device__ __forceinline__ void step_sc(const T* x, T* u, T dt, T* xn, T s, T c) const {
        const T vx = x[3], vy = x[4], vh = x[5];
        const T D0 = vx >= T(0) ? Dpos[0] : Dneg[0];
        const T D1 = vy >= T(0) ? Dpos[1] : Dneg[1];
        const T D2 = vh >= T(0) ? Dpos[2] : Dneg[2];
        const T wx = c * vx - s * vy;               // R(h) v, world-frame velocity
        const T wy = s * vx + c * vy;
        if (use_rudder) {                           // intermediate:71-77, advanced:101-108
            // The reference wraps (course angle - heading): ang = atan2(R(h) v), then atan2(sin(ang - h), cos(ang - h)).
            // R(h) v is v turned by h, so that difference IS the body-frame course angle atan2(vy, vx) -- one
            // atan2 instead of atan2 + sincos + atan2 on the rollout's dependent chain.  float32 engines take the
            // short form (it differs from the long one by rounding only); float64 engines keep the reference's
            // literal sequence, and so does v = 0 (seed states, the no-reverse clamp), where the reference's
            // atan2(+-0, +-0) of the rotated zero vector makes the term -h or pi - h instead of 0.
            T wr;
            if (sizeof(T) == 4 && !(vx == T(0) && vy == T(0))) {
                wr = atan2_fast(vy, vx);
            } else {
                const T ang = atan2_t(wy, wx);
                T sg, cg;
                sincos_t(ang, &sg, &cg);
                wr = wrap_sc(s, c, sg, cg);
            }
            u[2] = u[2] + rudder * wr;
        }
        if (thruster_clip) {                        // advanced:111  u = B clip(invB u, +-tmax)
            T th[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                T t = invB[i * 3 + 0] * u[0] + invB[i * 3 + 1] * u[1] + invB[i * 3 + 2] * u[2];
                th[i] = min_t(max_t(t, -tmax[i]), tmax[i]);
            }
#pragma unroll
            for (int i = 0; i < 3; i++)
                u[i] = B[i * 4 + 0] * th[0] + B[i * 4 + 1] * th[1] + B[i * 4 + 2] * th[2] + B[i * 4 + 3] * th[3];
        } else {                                    // novice:67-69
#pragma unroll
            for (int i = 0; i < 3; i++)
                if (abs_t(u[i]) > umax[i]) u[i] = u[i] > T(0) ? umax[i] : -umax[i];
        }
        xn[0] = x[0] + wx * dt;
        xn[1] = x[1] + wy * dt;
        xn[2] = x[2] + vh * dt;
        xn[3] = vx + (invM[0] * (u[0] - D0 * vx)) * dt;
        xn[4] = vy + (invM[1] * (u[1] - D1 * vy)) * dt;
        xn[5] = vh + (invM[2] * (u[2] - D2 * vh)) * dt;
        if (carlike) {                              // intermediate:88-98, advanced:120-128
            // float32: |xn3 / vmax| as a product with the host-rounded reciprocal (one FMUL instead of the 10-instruction
            // IEEE division on the chain; differs from the quotient by 1 ulp at most); float64 divides like the reference
            if (sizeof(T) == 4) {
                if (vx > T(0)) xn[5] = min_t(abs_t(xn[3] * inv_vpos0), T(1)) * xn[5];
                else if (vx < T(0)) xn[5] = min_t(abs_t(xn[3] * inv_vneg0), T(1)) * xn[5];
            } else {
                if (vx > T(0)) xn[5] = min_t(abs_t(xn[3] / vpos0), T(1)) * xn[5];
                else if (vx < T(0)) xn[5] = min_t(abs_t(xn[3] / vneg0), T(1)) * xn[5];
            }
            if (xn[3] < T(0)) xn[3] = T(0);
        }
    }
};

template <typename T> struct CarSys {             // demos/demo_car.py:46-72
    static constexpr int N = 5, M = 2;
    static constexpr int KIND = LQ_SYS_CAR;
    T invM[2], D[2], ulo[2], uhi[2], velmax0;
    static constexpr int NPARAMS = 9;
    __host__ void unpack(const double* p) {
        int k = 0;
        for (int i = 0; i < 2; i++) invM[i] = (T)p[k++];
        for (int i = 0; i < 2; i++) D[i] = (T)p[k++];
        for (int i = 0; i < 2; i++) ulo[i] = (T)p[k++];
        for (int i = 0; i < 2; i++) uhi[i] = (T)p[k++];
        velmax0 = (T)p[k++];
    }
    static constexpr int HEADING = 2;
    __device__ __forceinline__ void step(const T* x, T* u, T dt, T* xn) const {
        T s, c;
        sincos_t(x[2], &s, &c);
        step_sc(x, u, dt, xn, s, c);
    }
    __device__ __forceinline__ void step_sc(const T* x, T* u, T dt, T* xn, T s, T c) const {
        u[0] = min_t(max_t(u[0], ulo[0]), uhi[0]);
        u[1] = min_t(max_t(u[1], ulo[1]), uhi[1]);
        xn[0] = x[0] + (c * x[3]) * dt;
        xn[1] = x[1] + (s * x[3]) * dt;
        xn[2] = x[2] + x[4] * dt;
        xn[3] = x[3] + (invM[0] * (u[0] - D[0] * x[3])) * dt;
        xn[4] = x[4] + (invM[1] * (u[1] - D[1] * x[4])) * dt;
        if (xn[3] < T(0)) xn[3] = T(0);
        xn[4] = min_t(abs_t(xn[3] / velmax0), T(1)) * xn[4];
    }
};

template <typename T> struct PendulumSys {        // config C1
    static constexpr int N = 2, M = 1;
    static constexpr int KIND = LQ_SYS_PENDULUM2;
    T ml2, mgl, damping, tau_max;
    static constexpr int NPARAMS = 5;
    __host__ void unpack(const double* p) {
        ml2 = (T)(p[0] * p[1] * p[1]); mgl = (T)(p[0] * p[3] * p[1]); damping = (T)p[2]; tau_max = (T)p[4];
    }
    static constexpr int HEADING = 0;
    __device__ __forceinline__ void step(const T* x, T* u, T dt, T* xn) const {
        T s, c;
        sincos_t(x[0], &s, &c);
        step_sc(x, u, dt, xn, s, c);
    }
    __device__ __forceinline__ void step_sc(const T* x, T* u, T dt, T* xn, T s, T) const {
        const T tau = min_t(max_t(u[0], -tau_max), tau_max);
        const T wdot = (tau - damping * x[1] - mgl * s) / ml2;
        xn[0] = x[0] + x[1] * dt;
        xn[1] = x[1] + wdot * dt;
    }
};

template <typename T, int D> struct DoubleIntSys { // config C5
    static constexpr int N = 2 * D, M = D;
    static constexpr int KIND = LQ_SYS_DOUBLE_INTEGRATOR;
    int axes;
    T a_max;
    static constexpr int NPARAMS = 2;
    __host__ void unpack(const double* p) { axes = (int)p[0]; a_max = (T)p[1]; }
    static constexpr int HEADING = -1;
    __device__ __forceinline__ void step_sc(const T* x, T* u, T dt, T* xn, T, T) const { step(x, u, dt, xn); }
    __device__ __forceinline__ void step(const T* x, T* u, T dt, T* xn) const {
#pragma unroll
        for (int i = 0; i < D; i++) {
            u[i] = min_t(max_t(u[i], -a_max), a_max);
            xn[i] = x[i] + x[D + i] * dt;
            xn[D + i] = x[D + i] + u[i] * dt;
        }
    }
};

// ---------------------------------------------------------------------------
// erf (reference a7): e = xgoal - x, masked dims wrapped.
// ---------------------------------------------------------------------------
template <typename T, int N>
__device__ __forceinline__ void erf_eval(uint32_t angle_mask, const T* xgoal, const T* x, T* e) {
#pragma unroll
    for (int i = 0; i < N; i++) {
        if (angle_mask & (1u << i)) {
            T s, c, sg, cg;
            sincos_t(x[i], &s, &c);
            sincos_t(xgoal[i], &sg, &cg);
            e[i] = wrap_sc(s, c, sg, cg);
        } else {
            e[i] = xgoal[i] - x[i];
        }
    }
}

// Same, with sin/cos of dim HD of both states supplied by the caller (the rollout caches them per state).  The common
// case -- no wrapped dim, or only HD -- is straight-line code; other wrapped dims take the general path out of line,
// so their sin/cos/atan2 code is not replicated N times into the rollout loop.
template <typename T, int N>
__device__ __noinline__ void erf_wrap_others(uint32_t rest_mask, const T* xgoal, const T* x, T* e) {
    for (int i = 0; i < N; i++) {
        if (rest_mask & (1u << i)) {
            T s2, c2, sg2, cg2;
            sincos_t(x[i], &s2, &c2);
            sincos_t(xgoal[i], &sg2, &cg2);
            e[i] = wrap_sc(s2, c2, sg2, cg2);
        }
    }
}
template <typename T, int N, int HD>
__device__ __forceinline__ void erf_eval_sc(uint32_t angle_mask, const T* xgoal, const T* x, T* e, T sg, T cg, T s, T c) {
#pragma unroll
    for (int i = 0; i < N; i++) e[i] = xgoal[i] - x[i];
    constexpr uint32_t hbit = HD >= 0 ? (1u << (HD < 0 ? 0 : HD)) : 0u;
    if (HD >= 0 && (angle_mask & hbit)) e[HD < 0 ? 0 : HD] = wrap_sc(s, c, sg, cg);
    if (angle_mask & ~hbit) erf_wrap_others<T, N>(angle_mask & ~hbit, xgoal, x, e);
}

// ---------------------------------------------------------------------------
// lqr providers (reference a6).  gain(): K (M x N row-major) at (x,u);
// metric(): S (N x N row-major) at x with u = 0 (planner.py:348).
// ---------------------------------------------------------------------------
template <typename T, class Sys> struct PdRotatedLqr {
    static constexpr int N = Sys::N, M = Sys::M;
    static constexpr int KIND = LQ_LQR_PD_ROTATED;
    static constexpr bool CONST_METRIC = true;
    static constexpr bool COOP = false;
    static constexpr int COOP_SCALARS = 0;
    static constexpr int SLOT_SCALARS = 0;
    T kp[M * M], kd[M * M], S[N * N];
    __host__ int unpack(const double* p, int count) {
        if (count != 2 * M * M + N * N) return -1;
        if (!((N == 6 && M == 3) || (N == 5 && M == 2))) return -1;
        int k = 0;
        for (int i = 0; i < M * M; i++) kp[i] = (T)p[k++];
        for (int i = 0; i < M * M; i++) kd[i] = (T)p[k++];
        for (int i = 0; i < N * N; i++) S[i] = (T)p[k++];
        return 0;
    }
    __device__ __forceinline__ void gain(const Sys& sys, T dt, const T* x, const T* u, T* K) const {
        T s, c;
        sincos_t(x[2], &s, &c);
        gain_sc(sys, dt, x, u, K, s, c);
    }
    __device__ __forceinline__ void gain_sc(const Sys&, T, const T*, const T*, T* K, T s, T c) const {
        if (M == 3) {                               // K = [kp R(h)^T, kd], demo_boat_novice.py:86-103
#pragma unroll
            for (int i = 0; i < M; i++) {
                K[i * N + 0] = kp[i * M + 0] * c - kp[i * M + 1] * s;
                K[i * N + 1] = kp[i * M + 0] * s + kp[i * M + 1] * c;
                K[i * N + 2] = kp[i * M + 2 % M];
#pragma unroll
                for (int j = 0; j < M; j++) K[i * N + 3 + j] = kd[i * M + j];
            }
        } else {                                    // K = [kp [[c,s,0],[0,0,1]], kd], demo_car.py:95-113
#pragma unroll
            for (int i = 0; i < M; i++) {
                K[i * N + 0] = kp[i * M + 0] * c;
                K[i * N + 1] = kp[i * M + 0] * s;
                K[i * N + 2] = kp[i * M + 1 % M];
#pragma unroll
                for (int j = 0; j < M; j++) K[i * N + 3 + j] = kd[i * M + j];
            }
        }
    }
    __device__ __forceinline__ void metric(const Sys&, T, const T*, const T*, T* Sout) const {
#pragma unroll
        for (int i = 0; i < N * N; i++) Sout[i] = S[i];
    }
};

template <typename T, class Sys> struct ConstantLqr {
    static constexpr int N = Sys::N, M = Sys::M;
    static constexpr int KIND = LQ_LQR_CONSTANT;
    static constexpr bool CONST_METRIC = true;
    static constexpr bool COOP = false;
    static constexpr int COOP_SCALARS = 0;
    static constexpr int SLOT_SCALARS = 0;
    T S[N * N], K0[M * N];
    __host__ int unpack(const double* p, int count) {
        if (count != N * N + M * N) return -1;
        int k = 0;
        for (int i = 0; i < N * N; i++) S[i] = (T)p[k++];
        for (int i = 0; i < M * N; i++) K0[i] = (T)p[k++];
        return 0;
    }
    __device__ __forceinline__ void gain(const Sys&, T, const T*, const T*, T* K) const {
#pragma unroll
        for (int i = 0; i < M * N; i++) K[i] = K0[i];
    }
    __device__ __forceinline__ void gain_sc(const Sys& sys, T dt, const T* x, const T* u, T* K, T, T) const { gain(sys, dt, x, u, K); }
    __device__ __forceinline__ void metric(const Sys&, T, const T*, const T*, T* Sout) const {
#pragma unroll
        for (int i = 0; i < N * N; i++) Sout[i] = S[i];
    }
};

// Finite-difference Jacobian + backward finite-horizon discrete Riccati sweep
// (BASELINE.json north_star; oracle twin oracle/models.py fd_jacobians /
// riccati_sweep).  All matrices are thread-private; loops are fully unrolled
// for small N so they live in registers, larger N spills to L1-backed local.
template <typename T, class Sys> struct FdRiccatiLqr {
    static constexpr int N = Sys::N, M = Sys::M;
    static constexpr int KIND = LQ_LQR_FD_RICCATI;
    static constexpr bool CONST_METRIC = false;
    // Cooperative form: the L lanes that share one sample keep the matrices of the sweep in a slice of
    // shared memory and split every product by output element; finite-difference columns are split the
    // same way.  Each output is accumulated in ascending k exactly as in solve(), so both forms agree.
    static constexpr bool COOP = true;
    // one solve: A, B, P, BtP, G, Rhs/K, AmBK, AtP.  A group's slice holds COOP_SLOTS of them: the solve in progress and
    // the one before, so that a rollout whose Jacobian alternates between two values (an axis that saturates every
    // other step: saturated -> its gain row is zero -> no effort -> unsaturated -> ...) runs each sweep once.
    static constexpr int SLOT_SCALARS = N * N + N * M + N * N + M * N + M * M + M * N + N * N + N * N;
    static constexpr int COOP_SLOTS = 2;
    static constexpr int COOP_SCALARS = SLOT_SCALARS * COOP_SLOTS;
    T Q[N * N], R[M * M];
    T eps;
    int steps;
    // snap > 0: every finite-difference Jacobian entry is rounded to the nearest multiple of `snap` before the sweep.
    // For dynamics that are (piecewise) linear with entries on that grid this removes the finite-difference noise, and
    // the gain becomes a function of a handful of distinct Jacobians -- which is what lets a rollout keep its gain
    // while the Jacobian does not change (gain_coop below).  0 = the Jacobian as differenced.
    T snap;
    __host__ int unpack(const double* p, int count) {
        if (count != N * N + M * M + 2 && count != N * N + M * M + 3) return -1;
        int k = 0;
        for (int i = 0; i < N * N; i++) Q[i] = (T)p[k++];
        for (int i = 0; i < M * M; i++) R[i] = (T)p[k++];
        eps = (T)p[k++]; steps = (int)p[k++];
        snap = count == N * N + M * M + 3 ? (T)p[k++] : T(0);
        return 0;
    }
    __device__ __forceinline__ T snapped(T v) const { return snap > T(0) ? rint_t(v / snap) * snap : v; }

    __device__ void solve(const Sys& sys, T dt, const T* x, const T* u, T* P, T* K) const {
        T A[N * N], Bm[N * M];
        const T inv2e = T(1) / (T(2) * eps);
        // central differences of the discrete map (n + m column pairs)
        for (int j = 0; j < N; j++) {
            T xp[N], xm[N], up[M], um[M], fp[N], fm[N];
#pragma unroll
            for (int i = 0; i < N; i++) { xp[i] = x[i]; xm[i] = x[i]; }
#pragma unroll
            for (int i = 0; i < M; i++) { up[i] = u[i]; um[i] = u[i]; }
            xp[j] += eps; xm[j] -= eps;
            sys.step(xp, up, dt, fp);
            sys.step(xm, um, dt, fm);
#pragma unroll
            for (int i = 0; i < N; i++) A[i * N + j] = snapped((fp[i] - fm[i]) * inv2e);
        }
        for (int j = 0; j < M; j++) {
            T xp[N], xm[N], up[M], um[M], fp[N], fm[N];
#pragma unroll
            for (int i = 0; i < N; i++) { xp[i] = x[i]; xm[i] = x[i]; }
#pragma unroll
            for (int i = 0; i < M; i++) { up[i] = u[i]; um[i] = u[i]; }
            up[j] += eps; um[j] -= eps;
            sys.step(xp, up, dt, fp);
            sys.step(xm, um, dt, fm);
#pragma unroll
            for (int i = 0; i < N; i++) Bm[i * M + j] = snapped((fp[i] - fm[i]) * inv2e);
        }
#pragma unroll
        for (int i = 0; i < N * N; i++) P[i] = Q[i];
#pragma unroll
        for (int i = 0; i < M * N; i++) K[i] = T(0);
        for (int it = 0; it < steps; it++) {
            T BtP[M * N];                           // B' P        (M x N)
            for (int i = 0; i < M; i++)
                for (int j = 0; j < N; j++) {
                    T a = T(0);
                    for (int k = 0; k < N; k++) a += Bm[k * M + i] * P[k * N + j];
                    BtP[i * N + j] = a;
                }
            T G[M * M], Rhs[M * N];                 // G = R + B'PB, Rhs = B'PA
            for (int i = 0; i < M; i++) {
                for (int j = 0; j < M; j++) {
                    T a = R[i * M + j];
                    for (int k = 0; k < N; k++) a += BtP[i * N + k] * Bm[k * M + j];
                    G[i * M + j] = a;
                }
                for (int j = 0; j < N; j++) {
                    T a = T(0);
                    for (int k = 0; k < N; k++) a += BtP[i * N + k] * A[k * N + j];
                    Rhs[i * N + j] = a;
                }
            }
            // K = G^-1 Rhs by Gaussian elimination with partial pivoting (as LAPACK gesv in the oracle)
            for (int col = 0; col < M; col++) {
                int piv = col;
                T big = abs_t(G[col * M + col]);
                for (int r = col + 1; r < M; r++) {
                    T v = abs_t(G[r * M + col]);
                    if (v > big) { big = v; piv = r; }
                }
                if (piv != col) {
                    for (int j = 0; j < M; j++) { T t = G[col * M + j]; G[col * M + j] = G[piv * M + j]; G[piv * M + j] = t; }
                    for (int j = 0; j < N; j++) { T t = Rhs[col * N + j]; Rhs[col * N + j] = Rhs[piv * N + j]; Rhs[piv * N + j] = t; }
                }
                const T inv = T(1) / G[col * M + col];
                for (int r = col + 1; r < M; r++) {
                    const T f = G[r * M + col] * inv;
                    for (int j = col; j < M; j++) G[r * M + j] -= f * G[col * M + j];
                    for (int j = 0; j < N; j++) Rhs[r * N + j] -= f * Rhs[col * N + j];
                }
            }
            for (int r = M - 1; r >= 0; r--) {
                const T inv = T(1) / G[r * M + r];
                for (int j = 0; j < N; j++) {
                    T a = Rhs[r * N + j];
                    for (int k = r + 1; k < M; k++) a -= G[r * M + k] * K[k * N + j];
                    K[r * N + j] = a * inv;
                }
            }
            // P = Q + A' P (A - B K)
            T AmBK[N * N];
            for (int i = 0; i < N; i++)
                for (int j = 0; j < N; j++) {
                    T a = A[i * N + j];
                    for (int k = 0; k < M; k++) a -= Bm[i * M + k] * K[k * N + j];
                    AmBK[i * N + j] = a;
                }
            T AtP[N * N];
            for (int i = 0; i < N; i++)
                for (int j = 0; j < N; j++) {
                    T a = T(0);
                    for (int k = 0; k < N; k++) a += A[k * N + i] * P[k * N + j];
                    AtP[i * N + j] = a;
                }
            for (int i = 0; i < N; i++)
                for (int j = 0; j < N; j++) {
                    T a = Q[i * N + j];
                    for (int k = 0; k < N; k++) a += AtP[i * N + k] * AmBK[k * N + j];
                    P[i * N + j] = a;
                }
        }
    }
    // All `lanes` threads of the group call this with identical (x, u); on return sm holds P (offset
    // P_OFF) and K (offset K_OFF).  sync() must order shared-memory accesses within the group.
    static constexpr int A_OFF = 0, B_OFF = A_OFF + N * N, P_OFF = B_OFF + N * M, BTP_OFF = P_OFF + N * N,
                         G_OFF = BTP_OFF + M * N, K_OFF = G_OFF + M * M, AMBK_OFF = K_OFF + M * N,
                         ATP_OFF = AMBK_OFF + N * N;
    // memo (may be null): the group's rollout-long note of which slots hold finished solves (bit s = slot s valid, bit 2 =
    // the slot used last).  If the new Jacobian equals a held one bit for bit, that slot's P and K are what the sweep
    // would produce again: it is skipped, the result is identical by construction.  Returns the slot that holds P and K.
    template <class Sync>
    __device__ int solve_coop(const Sys& sys, T dt, const T* x, const T* u, T* sm, int lane, int lanes, Sync sync,
                              int* memo = nullptr) const {
        const T inv2e = T(1) / (T(2) * eps);
        // finite-difference column j of [A | B] into v
        auto fd_column = [&](int j, T (&v)[N]) {
            T xp[N], xm[N], up[M], um[M], fp[N], fm[N];
#pragma unroll
            for (int i = 0; i < N; i++) { xp[i] = x[i]; xm[i] = x[i]; }
#pragma unroll
            for (int i = 0; i < M; i++) { up[i] = u[i]; um[i] = u[i]; }
#pragma unroll
            for (int i = 0; i < N; i++) if (i == j) { xp[i] += eps; xm[i] -= eps; }
#pragma unroll
            for (int i = 0; i < M; i++) if (N + i == j) { up[i] += eps; um[i] -= eps; }
            sys.step(xp, up, dt, fp);
            sys.step(xm, um, dt, fm);
#pragma unroll
            for (int i = 0; i < N; i++) v[i] = snapped((fp[i] - fm[i]) * inv2e);
        };
        auto put_column = [&](T* slot, int j, const T (&v)[N]) {
            if (j < N) {
#pragma unroll
                for (int i = 0; i < N; i++) slot[A_OFF + i * N + j] = v[i];
            } else {
#pragma unroll
                for (int i = 0; i < N; i++) slot[B_OFF + i * M + (j - N)] = v[i];
            }
        };
        int slot = 0;
        sync();                                        // previous users of the slice are done
        if (memo && lanes >= N + M) {                  // one column per lane: compare before anything is overwritten
            const int state = *memo;
            T v[N];
            bool m0 = (state & 1) != 0, m1 = (state & 2) != 0;
            if (lane < N + M) {
                fd_column(lane, v);
#pragma unroll
                for (int i = 0; i < N; i++) {
                    const int off = lane < N ? A_OFF + i * N + lane : B_OFF + i * M + (lane - N);
                    if (m0) m0 = sm[off] == v[i];
                    if (m1) m1 = sm[SLOT_SCALARS + off] == v[i];
                }
            }
            m0 = sync.all(m0); m1 = sync.all(m1);      // every lane of the group takes the same branch
            if (m0) { *memo = (state & 3); return 0; }
            if (m1) { *memo = (state & 3) | 4; return 1; }
            slot = !(state & 1) ? 0 : (!(state & 2) ? 1 : ((state & 4) ? 0 : 1));     // an empty slot, else not the last used
            *memo = ((state & 3) | (1 << slot)) & 3 | (slot << 2);
            sm += slot * SLOT_SCALARS;
            if (lane < N + M) put_column(sm, lane, v);
        } else {
            for (int j = lane; j < N + M; j += lanes) {
                T v[N];
                fd_column(j, v);
                put_column(sm, j, v);
            }
        }
        T* A = sm + A_OFF; T* Bm = sm + B_OFF; T* P = sm + P_OFF; T* BtP = sm + BTP_OFF;
        T* G = sm + G_OFF; T* K = sm + K_OFF; T* AmBK = sm + AMBK_OFF; T* AtP = sm + ATP_OFF;
        for (int e = lane; e < N * N; e += lanes) P[e] = Q[e];
        sync();
        for (int it = 0; it < steps; it++) {
            for (int e = lane; e < M * N + N * N; e += lanes) {
                if (e < M * N) {                       // BtP = B' P
                    const int i = e / N, j = e % N;
                    T a = T(0);
                    for (int k = 0; k < N; k++) a += Bm[k * M + i] * P[k * N + j];
                    BtP[e] = a;
                } else {                               // AtP = A' P
                    const int q = e - M * N, i = q / N, j = q % N;
                    T a = T(0);
                    for (int k = 0; k < N; k++) a += A[k * N + i] * P[k * N + j];
                    AtP[q] = a;
                }
            }
            sync();
            for (int e = lane; e < M * M + M * N; e += lanes) {
                if (e < M * M) {                       // G = R + B'P B
                    const int i = e / M, j = e % M;
                    T a = R[e];
                    for (int k = 0; k < N; k++) a += BtP[i * N + k] * Bm[k * M + j];
                    G[e] = a;
                } else {                               // Rhs = B'P A (stored where K will be)
                    const int q = e - M * M, i = q / N, j = q % N;
                    T a = T(0);
                    for (int k = 0; k < N; k++) a += BtP[i * N + k] * A[k * N + j];
                    K[q] = a;
                }
            }
            sync();
            {   // K = G^-1 Rhs: every lane factors its own copy of the M x M system (identical pivots) and
                // carries the right-hand-side columns it owns through it.  Fully unrolled with predicated
                // row swaps so Gl / mult / rhs stay in registers (no dynamically indexed local arrays).
                T Gl[M * M], mult[M * M];
                int perm[M];
#pragma unroll
                for (int e = 0; e < M * M; e++) { Gl[e] = G[e]; mult[e] = T(0); }
#pragma unroll
                for (int col = 0; col < M; col++) {
                    int piv = col;
                    T big = abs_t(Gl[col * M + col]);
#pragma unroll
                    for (int r = col + 1; r < M; r++) {
                        const T v = abs_t(Gl[r * M + col]);
                        if (v > big) { big = v; piv = r; }
                    }
                    perm[col] = piv;
#pragma unroll
                    for (int r = col + 1; r < M; r++) {
                        if (r == piv) {
#pragma unroll
                            for (int j = 0; j < M; j++) { const T t = Gl[col * M + j]; Gl[col * M + j] = Gl[r * M + j]; Gl[r * M + j] = t; }
                        }
                    }
                    const T inv = T(1) / Gl[col * M + col];
#pragma unroll
                    for (int r = col + 1; r < M; r++) {
                        const T f = Gl[r * M + col] * inv;
                        mult[r * M + col] = f;
#pragma unroll
                        for (int j = col; j < M; j++) Gl[r * M + j] -= f * Gl[col * M + j];
                    }
                }
                for (int j = lane; j < N; j += lanes) {
                    T rhs[M];
#pragma unroll
                    for (int r = 0; r < M; r++) rhs[r] = K[r * N + j];
#pragma unroll
                    for (int col = 0; col < M; col++) {
#pragma unroll
                        for (int r = col + 1; r < M; r++)
                            if (r == perm[col]) { const T t = rhs[col]; rhs[col] = rhs[r]; rhs[r] = t; }
#pragma unroll
                        for (int r = col + 1; r < M; r++) rhs[r] -= mult[r * M + col] * rhs[col];
                    }
#pragma unroll
                    for (int r = M - 1; r >= 0; r--) {
                        T a = rhs[r];
#pragma unroll
                        for (int k = r + 1; k < M; k++) a -= Gl[r * M + k] * rhs[k];
                        rhs[r] = a * (T(1) / Gl[r * M + r]);
                    }
#pragma unroll
                    for (int r = 0; r < M; r++) K[r * N + j] = rhs[r];
                }
            }
            sync();
            for (int e = lane; e < N * N; e += lanes) {   // A - B K
                const int i = e / N, j = e % N;
                T a = A[e];
                for (int k = 0; k < M; k++) a -= Bm[i * M + k] * K[k * N + j];
                AmBK[e] = a;
            }
            sync();
            for (int e = lane; e < N * N; e += lanes) {   // P = Q + A'P (A - B K)
                const int i = e / N, j = e % N;
                T a = Q[e];
                for (int k = 0; k < N; k++) a += AtP[i * N + k] * AmBK[k * N + j];
                P[e] = a;
            }
            sync();
        }
        return slot;
    }
    template <class Sync>
    __device__ void gain_coop(const Sys& sys, T dt, const T* x, const T* u, T* K, T* sm, int lane, int lanes, Sync sync,
                              int* memo = nullptr) const {
        const int slot = solve_coop(sys, dt, x, u, sm, lane, lanes, sync, memo);
#pragma unroll
        for (int e = 0; e < M * N; e++) K[e] = sm[slot * SLOT_SCALARS + K_OFF + e];
    }

    __device__ void gain(const Sys& sys, T dt, const T* x, const T* u, T* K) const {
        T P[N * N];
        solve(sys, dt, x, u, P, K);
    }
    __device__ void gain_sc(const Sys& sys, T dt, const T* x, const T* u, T* K, T, T) const { gain(sys, dt, x, u, K); }
    __device__ void metric(const Sys& sys, T dt, const T* x, const T* u, T* Sout) const {
        T K[M * N];
        solve(sys, dt, x, u, Sout, K);
    }
};

// ---------------------------------------------------------------------------
// feasibility (reference a8): conjunction of primitives, obstacle tables in
// global memory (read-only path; every lane of a warp walks the same table).
// ---------------------------------------------------------------------------
template <typename T> struct FeasPrim {
    int kind, n_obs, n_verts, flags;
    int dims[4];
    T inflate;
    T bound_r;            // cloud: max vertex norm (broad phase)
    const T* obs;         // n_obs x 4 (row-major)
    const T* verts;       // 2 x n_verts
    T lo[LQ_MAX_N], hi[LQ_MAX_N];
    int nbox;
    // Uniform-grid broad phase over the obstacle table (host-built, engine.cuh build_grid): cell c lists
    // every obstacle whose zone of influence (its bounds grown by the hull's bounding radius, plus a
    // margin) overlaps the cell, so a query only walks the list of the cell it falls in.  Exact: an
    // obstacle that can be hit from p always has p inside its zone.  cell_start == nullptr => brute force.
    const int* cell_start;
    const int* cell_items;
    int n_cells, n_items;  // entries of cell_start (minus one) / cell_items
    T g0[3], ginv[3];
    int gn[3];
    // Product-grid hull (every demo's vertex cloud is np.mgrid, demo_car.py:77-90): verts = {hx_i} x {hy_j}.
    // The squared distance from an obstacle centre to its nearest vertex then separates per body axis,
    // min_i (bx - hx_i)^2 + min_j (by - hy_j)^2, so "any vertex within r" costs nx + ny tests, not nx * ny.
    const T* hull_axes;   // [nx][ny] values, nullptr when the cloud is not a product grid
    int hull_nx, hull_ny;
};

template <typename T> struct FeasSet {
    int n_prims;
    FeasPrim<T> prims[LQ_MAX_PRIMS];
};

// Where a primitive's tables are read from.  The rollout kernel copies the tables of a small obstacle field
// (grid cell lists, obstacle rows, hull axes: 18 KB for C4's 500 circles) into shared memory once per CTA, so
// the 3-4 dependent loads of every feasibility test cost shared-memory latency (29 cycles) instead of L2
// latency (~250 cycles each) on the rollout's dependent chain; kernels that do not stage read global memory.
template <typename T> struct FeasTab {
    const T* obs;
    const int* cell_start;
    const int* cell_items;
    const T* hull_axes;
};

// one obstacle row {a, b, c, d} with vector loads (rows are 16-byte aligned: cudaMalloc / 16-byte staged)
__device__ __forceinline__ void load_row4(const float* p, float& a, float& b, float& c, float& d) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    a = v.x; b = v.y; c = v.z; d = v.w;
}
__device__ __forceinline__ void load_row4(const double* p, double& a, double& b, double& c, double& d) {
    const double2 v = *reinterpret_cast<const double2*>(p), w = *reinterpret_cast<const double2*>(p + 2);
    a = v.x; b = v.y; c = w.x; d = w.y;
}

// Candidate obstacle range for a query position: [*b, *e) indexes cell_items (or the table itself
// when there is no grid).  Returns false if the position lies outside every zone of influence.
template <typename T>
__device__ __forceinline__ bool feas_candidates(const FeasPrim<T>& p, const int* cell_start, T px, T py, T pz, int nd, int* b, int* e) {
    if (!cell_start) { *b = 0; *e = p.n_obs; return true; }
    const T fx = (px - p.g0[0]) * p.ginv[0], fy = (py - p.g0[1]) * p.ginv[1];
    if (!(fx >= T(0) && fy >= T(0) && fx < (T)p.gn[0] && fy < (T)p.gn[1])) return false;
    int cell = (int)fy * p.gn[0] + (int)fx;
    if (nd == 3) {
        const T fz = (pz - p.g0[2]) * p.ginv[2];
        if (!(fz >= T(0) && fz < (T)p.gn[2])) return false;
        cell += (int)fz * p.gn[0] * p.gn[1];
    }
    *b = cell_start[cell]; *e = cell_start[cell + 1];
    return true;
}

// lane / lanes: cooperative evaluation by `lanes` threads that all hold the
// same (x,u); obstacle loops are strided by lane.  Returns this lane's verdict;
// the caller combines lanes (AND).  lanes == 1 is the scalar case.
// hdim >= 0: (hs, hc) = sincos(x[hdim]) is supplied and reused by the hull rotations on that dim.
// tabs: per-primitive table pointers (shared-memory copies in the rollout kernel); nullptr = the global tables.
template <typename T, int N, int M>
__device__ __forceinline__ bool feasible_eval(const FeasSet<T>& fs, const T* x, const T* u, int lane = 0, int lanes = 1,
                                              int hdim = -1, T hs = T(0), T hc = T(1), const FeasTab<T>* tabs = nullptr) {
    for (int pi = 0; pi < fs.n_prims; pi++) {
        const FeasPrim<T>& p = fs.prims[pi];
        switch (p.kind) {
        case LQ_FEAS_STATE_BOX: {                   // unused dims carry -inf / +inf bounds (engine.cuh init)
            bool out = false;
#pragma unroll
            for (int i = 0; i < N; i++) out = out || x[i] > p.hi[i] || x[i] < p.lo[i];
            if (out) return false;
            break;
        }
        case LQ_FEAS_EFFORT_BOX: {
            bool out = false;
#pragma unroll
            for (int i = 0; i < M; i++) out = out || abs_t(u[i]) > p.hi[i];
            if (out) return false;
            break;
        }
        case LQ_FEAS_STATE_REGION: {
            bool inside = true;
#pragma unroll
            for (int i = 0; i < N; i++)
                if (i < p.nbox && !(x[i] >= p.lo[i] && x[i] <= p.hi[i])) inside = false;
            if (inside) return false;
            break;
        }
        case LQ_FEAS_CIRCLES_POINT:                 // demo_boat_novice.py:160-164
        case LQ_FEAS_SPHERES_POINT: {
            const T* obs = tabs ? tabs[pi].obs : p.obs;
            const int* cs = tabs ? tabs[pi].cell_start : p.cell_start;
            const int* ci = tabs ? tabs[pi].cell_items : p.cell_items;
            const bool sph = p.kind == LQ_FEAS_SPHERES_POINT;
            T px = T(0), py = T(0), pz = T(0);
#pragma unroll
            for (int i = 0; i < N; i++) {
                if (i == p.dims[0]) px = x[i];
                if (i == p.dims[1]) py = x[i];
                if (sph && i == p.dims[2]) pz = x[i];
            }
            int b, e;
            if (!feas_candidates(p, cs, px, py, pz, sph ? 3 : 2, &b, &e)) break;
            for (int k = b + lane; k < e; k += lanes) {
                const int o = cs ? ci[k] : k;
                T ox, oy, oc, od;
                load_row4(obs + o * 4, ox, oy, oc, od);
                const T dx = px - ox, dy = py - oy;
                const T dz = sph ? pz - oc : T(0);
                const T rr = sph ? od : oc;                                   // inflate + r, host-added in float64
                if (dx * dx + dy * dy + dz * dz <= rr * rr) return false;
            }
            break;
        }
        case LQ_FEAS_CIRCLES_CLOUD:                 // demo_boat_intermediate.py:198-210, demo_car.py:168-180
        case LQ_FEAS_BOXES_CLOUD: {
            const T* obs = tabs ? tabs[pi].obs : p.obs;
            const int* cs = tabs ? tabs[pi].cell_start : p.cell_start;
            const int* ci = tabs ? tabs[pi].cell_items : p.cell_items;
            const T* hull = tabs ? tabs[pi].hull_axes : p.hull_axes;
            T px = T(0), py = T(0), ph = T(0);
#pragma unroll
            for (int i = 0; i < N; i++) {
                if (i == p.dims[0]) px = x[i];
                if (i == p.dims[1]) py = x[i];
                if (i == p.dims[2]) ph = x[i];
            }
            const bool circ = p.kind == LQ_FEAS_CIRCLES_CLOUD;
            int b, e;
            if (p.flags & 1) {                      // demo_car.py:175: an extra vertex at 2 * position
                const T qx = px + px, qy = py + py;
                if (feas_candidates(p, cs, qx, qy, T(0), 2, &b, &e)) {
                    for (int k = b + lane; k < e; k += lanes) {
                        const int o = cs ? ci[k] : k;
                        T ox, oy, a, bb;
                        load_row4(obs + o * 4, ox, oy, a, bb);
                        const T ex = qx - ox, ey = qy - oy;
                        if (circ ? (ex * ex + ey * ey <= a * a) : (abs_t(ex) <= a && abs_t(ey) <= bb)) return false;
                    }
                }
            }
            if (!feas_candidates(p, cs, px, py, T(0), 2, &b, &e)) break;
            T s = hs, c = hc;
            if (p.dims[2] != hdim) sincos_t(ph, &s, &c);
            if (circ && hull) {
                // lanes split the candidate list; each candidate is O(nx + ny)
                for (int k = b + lane; k < e; k += lanes) {
                    const int o = cs ? ci[k] : k;
                    T ox, oy, a, unused;
                    load_row4(obs + o * 4, ox, oy, a, unused);
                    const T rx = ox - px, ry = oy - py;                               // centre relative to the hull origin
                    const T reach = a + p.bound_r;
                    if (rx * rx + ry * ry > reach * reach * T(1.0001)) continue;
                    const T bx = c * rx + s * ry, by = c * ry - s * rx;                // ... in the body frame
                    T mx = Num<T>::inf(), my = Num<T>::inf();
                    for (int i = 0; i < p.hull_nx; i++) { const T d = bx - hull[i]; mx = min_t(mx, d * d); }
                    for (int j = 0; j < p.hull_ny; j++) { const T d = by - hull[p.hull_nx + j]; my = min_t(my, d * d); }
                    if (mx + my <= a * a) return false;
                }
                break;
            }
            // general cloud: every lane walks the (short) candidate list, the vertices are what the lanes split
            for (int k = b; k < e; k++) {
                const int o = cs ? ci[k] : k;
                T ox, oy, a, bb;
                load_row4(obs + o * 4, ox, oy, a, bb);
                // hull bounding circle vs obstacle bounding circle (conservative)
                const T dx = px - ox, dy = py - oy;
                const T reach = (circ ? a : sqrt_t(a * a + bb * bb)) + p.bound_r;
                if (dx * dx + dy * dy > reach * reach * T(1.0001)) continue;
                for (int v = lane; v < p.n_verts; v += lanes) {
                    const T bx = p.verts[v], by = p.verts[p.n_verts + v];
                    const T ex = dx + (c * bx - s * by), ey = dy + (s * bx + c * by);
                    if (circ ? (ex * ex + ey * ey <= a * a) : (abs_t(ex) <= a && abs_t(ey) <= bb)) return false;
                }
            }
            break;
        }
        case LQ_FEAS_OGRID_CLOUD: {                 // demos/lqrrt_ros/nodes/lqrrt_node.py:730-745
            T px = T(0), py = T(0), ph = T(0);
#pragma unroll
            for (int i = 0; i < N; i++) {
                if (i == p.dims[0]) px = x[i];
                if (i == p.dims[1]) py = x[i];
                if (i == p.dims[2]) ph = x[i];
            }
            T s = hs, c = hc;
            if (p.dims[2] != hdim) sincos_t(ph, &s, &c);
            const long long rows = p.n_obs, cols = (long long)p.lo[0];
            const T cpm = p.lo[1], ox = p.lo[2], oy = p.lo[3], thr = p.lo[4];
            for (int v = lane; v < p.n_verts; v += lanes) {
                const T bx = p.verts[v], by = p.verts[p.n_verts + v];
                const T wx = px + (c * bx - s * by), wy = py + (s * bx + c * by);
                long long ix = (long long)(cpm * (wx - ox)), iy = (long long)(cpm * (wy - oy));   // astype(int64): toward zero
                if (ix < 0) ix += cols;             // NumPy wraps negative indices once ...
                if (iy < 0) iy += rows;
                if (ix < 0 || ix >= cols || iy < 0 || iy >= rows) return false;                   // ... IndexError => infeasible
                if (!(p.obs[iy * cols + ix] < thr)) return false;
            }
            break;
        }
        default: break;
        }
    }
    return true;
}

// ---- shared-memory staging of the obstacle tables (rollout kernels) --------------------------------------
// The engine keeps one contiguous IMAGE of all staged tables in global memory, laid out exactly as the kernel
// wants them in shared memory (every table 16-byte aligned; host-side layout in engine.cuh build_stage_image).
// A CTA brings it in with ONE bulk asynchronous copy (TMA: cp.async.bulk.shared.global, completion counted on
// an mbarrier), i.e. one L2 round trip for all ~18 KB instead of a dependent load -> store chain per element.
struct StagePlan {
    int on[LQ_MAX_PRIMS];
    int off_obs[LQ_MAX_PRIMS], off_cs[LQ_MAX_PRIMS], off_ci[LQ_MAX_PRIMS], off_hull[LQ_MAX_PRIMS];   // byte offsets, -1 = none
    int bytes;                 // image size = dynamic shared memory the staged tables need (multiple of 16)
    const unsigned char* image;
};

__host__ __device__ inline int stage_align(int b) { return (b + 15) & ~15; }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src), "r"(bytes),
                   "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned done = 0;
    const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    }
}

// Called by every thread of the CTA; `tabs` (LQ_MAX_PRIMS entries), `bar` and `raw` (16-byte aligned, sp.bytes
// long) live in shared memory.  Returns once the tables are readable by the whole CTA.
template <typename T>
__device__ __forceinline__ void stage_tables(const FeasSet<T>& fs, const StagePlan& sp, FeasTab<T>* tabs, unsigned char* raw,
                                             unsigned long long* bar) {
    if (threadIdx.x == 0) {
        if (sp.bytes > 0) {
            mbar_init(bar, 1);
            mbar_expect_tx(bar, (unsigned)sp.bytes);
            bulk_copy_g2s(raw, sp.image, (unsigned)sp.bytes, bar);
        }
        for (int pi = 0; pi < fs.n_prims; pi++) {
            const FeasPrim<T>& p = fs.prims[pi];
            FeasTab<T> t{p.obs, p.cell_start, p.cell_items, p.hull_axes};
            if (sp.on[pi]) {
                if (sp.off_obs[pi] >= 0) t.obs = reinterpret_cast<const T*>(raw + sp.off_obs[pi]);
                if (sp.off_cs[pi] >= 0) t.cell_start = reinterpret_cast<const int*>(raw + sp.off_cs[pi]);
                if (sp.off_ci[pi] >= 0) t.cell_items = reinterpret_cast<const int*>(raw + sp.off_ci[pi]);
                if (sp.off_hull[pi] >= 0) t.hull_axes = reinterpret_cast<const T*>(raw + sp.off_hull[pi]);
            }
            tabs[pi] = t;
        }
    }
    __syncthreads();                       // barrier initialised, table pointers published
    if (sp.bytes > 0) mbar_wait(bar, 0);   // ... and the image has landed
}

}  // namespace lq
