// pb_selftest_arith: the branch-deferred divide / reciprocal / square root of ArithFast
// (pb_device.cuh) against the compiler's IEEE a/b, 1.0/x, sqrt(x) on pseudo-random operand pairs
// spread over a wide exponent range, plus the special values the hot path meets (+-0 numerators).
// counts[0..3] = mismatching results of div, divz, rcp, root among the cases the fast path accepted;
// counts[4] = cases the fast path rejected (they take the exact path in the kernels).
#include "pb_device.cuh"
#include "pb_host.hpp"

namespace pb {
__device__ inline unsigned long long splitmix(unsigned long long &s) {
  unsigned long long z = (s += 0x9e3779b97f4a7c15ull);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}
__device__ inline Real rnd_real(unsigned long long &s, int span) {
  const unsigned long long u = splitmix(s);
  const Real m = 1.0 + (Real)(u >> 12) * (1.0 / 4503599627370496.0); // [1,2)
  const int e = (int)(splitmix(s) % (unsigned)(2 * span + 1)) - span;
  const Real v = ldexp(m, e);
  return (splitmix(s) & 1) ? -v : v;
}
__device__ inline bool same(Real a, Real b) { return __double_as_longlong(a) == __double_as_longlong(b) || (isnan(a) && isnan(b)); }

__global__ void arith_selftest_kernel(long long n, unsigned long long seed, int span, unsigned long long *counts) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= n) return;
  unsigned long long s = seed + 0x51ed270b7f1dull * (unsigned long long)t;
  Real a = rnd_real(s, span), b = rnd_real(s, span);
  const unsigned sel = (unsigned)(splitmix(s) & 15);
  if (sel >= 8) { // operands a few ulps (or a small power of two) away from a power of two: Lorentz factors of slow flows
    const int sh = (int)(splitmix(s) % 53);
    const Real eps = ldexp((Real)(splitmix(s) % 9), -sh - 1);
    b = (sel & 1) ? 1.0 - eps : 1.0 + eps;
    if (sel & 2) b = -b;
    if (sel & 4) a = 1.0 + ldexp((Real)(splitmix(s) % 9), -(int)(splitmix(s) % 53) - 1);
  }
  if (sel == 0) a = 0.0;
  if (sel == 1) a = -0.0;
  if (sel == 2) a = b;
  ArithExact ax;
  {
    ArithFast af;
    const Real q = af.div(a, b);
    if (af.failed()) atomicAdd(&counts[4], 1ull);
    else if (!same(q, ax.div(a, b))) atomicAdd(&counts[0], 1ull);
  }
  {
    ArithFast af;
    const Real q = af.divz(a, b);
    if (af.failed()) atomicAdd(&counts[4], 1ull);
    else if (!same(q, a / b)) atomicAdd(&counts[1], 1ull);
  }
  {
    ArithFast af;
    const ArithFast::Rcp d = af.prep(b);
    const Real q1 = af.divz(a, d), q2 = af.div(b, d);
    if (af.failed()) atomicAdd(&counts[4], 1ull);
    else if (!same(q1, a / b) || !same(q2, b / b)) atomicAdd(&counts[1], 1ull);
  }
  {
    ArithFast af;
    const Real r = af.rcp(b);
    if (af.failed()) atomicAdd(&counts[4], 1ull);
    else if (!same(r, 1.0 / b)) atomicAdd(&counts[2], 1ull);
  }
  {
    ArithFast af;
    const Real x = fabs(a);
    const Real r = af.root(x);
    if (af.failed()) atomicAdd(&counts[4], 1ull);
    else if (!same(r, sqrt(x))) atomicAdd(&counts[3], 1ull);
  }
}
} // namespace pb

extern "C" int pb_selftest_arith(long long n, unsigned long long seed, int exponent_span, unsigned long long *counts_dev,
                                 void *stream) {
  using namespace pb;
  if (n <= 0 || !counts_dev) return set_error(PB_EINVAL, "pb_selftest_arith: bad arguments");
  const int T = 256;
  arith_selftest_kernel<<<(unsigned)((n + T - 1) / T), T, 0, (cudaStream_t)stream>>>(n, seed, exponent_span, counts_dev);
  return check_launch("arith_selftest_kernel");
}

// ---- prim_to_flux with ArithFast against ArithExact on synthetic face states -------------------
namespace pb {
__global__ void riemann_selftest_kernel(long long n, unsigned long long seed, const pb_cfg_t cfg, int flat,
                                        unsigned long long *counts, double *first) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= n) return;
  unsigned long long s = seed + 0x9e3779b97f4a7c15ull * (unsigned long long)t;
  auto uni = [&](Real lo, Real hi) { return lo + (hi - lo) * ((Real)(splitmix(s) >> 11) * (1.0 / 9007199254740992.0)); };
  constexpr int NP = NC<true>::nprim, NCn = NC<true>::ncons;
  Real q[NP];
  const unsigned sel = (unsigned)(splitmix(s) & 7);
  q[PRHO] = (sel & 1) ? 1.0 : uni(1e-6, 10.0);
  for (int d = 0; d < 3; ++d) q[PVEL + d] = (sel & 2) ? 0.0 : uni(-2.0, 2.0);
  q[PENG] = uni(1e-8, 10.0);
  for (int d = 0; d < 3; ++d) q[PB_ + d] = (sel & 4) ? 0.0 : uni(-1.0, 1.0);
  q[NC<true>::prs] = cfg.gm1 * q[PENG];
  q[NC<true>::gm1] = cfg.gm1 + 1.0;
  PointGeom g;
  if (flat) {
    g.alpha = 1.0;
    g.detgam = 1.0;
    for (int d = 0; d < 3; ++d) g.beta[d] = 0.0;
    for (int m = 0; m < 10; ++m) g.gcov[m] = 0.0;
    g.gcov[flat4(0, 0)] = -1.0;
    for (int d = 1; d < 4; ++d) g.gcov[flat4(d, d)] = 1.0;
  } else {
    g.alpha = uni(0.3, 1.0);
    g.detgam = uni(0.5, 30.0);
    for (int d = 0; d < 3; ++d) g.beta[d] = uni(-0.3, 0.3);
    for (int m = 0; m < 10; ++m) g.gcov[m] = uni(-0.05, 0.05);
    g.gcov[flat4(0, 0)] = -uni(0.5, 1.0);
    for (int d = 1; d < 4; ++d) g.gcov[flat4(d, d)] = uni(1.0, 20.0);
  }
  const Real gdd = flat ? 1.0 : uni(0.05, 1.0);
  FloorCtx fc{1e-8, 1e-8, 1.0, 0.5, 0.5};
  Real Ua[NCn], Fa[NCn], Ub[NCn], Fb[NCn], vma, vpa, vmb, vpb;
  ArithFast af;
  ArithExact ax;
  prim_to_flux<1, true>(af, cfg, fc, g, gdd, q, vma, vpa, Ua, Fa);
  prim_to_flux<1, true>(ax, cfg, fc, g, gdd, q, vmb, vpb, Ub, Fb);
  if (af.failed()) {
    atomicAdd(&counts[1], 1ull);
    return;
  }
  int badc = -1;
  for (int m = 0; m < NCn; ++m) {
    if (!same(Ua[m], Ub[m])) badc = m;
    if (!same(Fa[m], Fb[m])) badc = 16 + m;
  }
  if (!same(vma, vmb)) badc = 32;
  if (!same(vpa, vpb)) badc = 33;
  if (badc >= 0 && atomicAdd(&counts[0], 1ull) == 0) {
    first[0] = badc;
    for (int m = 0; m < NP; ++m) first[1 + m] = q[m];
    first[12] = badc < 16 ? Ua[badc] : (badc < 32 ? Fa[badc - 16] : (badc == 32 ? vma : vpa));
    first[13] = badc < 16 ? Ub[badc] : (badc < 32 ? Fb[badc - 16] : (badc == 32 ? vmb : vpb));
  }
}
} // namespace pb

extern "C" int pb_selftest_riemann(long long n, unsigned long long seed, const pb_cfg_t *cfg, int flat,
                                   unsigned long long *counts_dev, double *first_dev, void *stream) {
  using namespace pb;
  if (n <= 0 || !counts_dev || !first_dev || !cfg) return set_error(PB_EINVAL, "pb_selftest_riemann: bad arguments");
  const int T = 128;
  riemann_selftest_kernel<<<(unsigned)((n + T - 1) / T), T, 0, (cudaStream_t)stream>>>(n, seed, *cfg, flat, counts_dev, first_dev);
  return check_launch("riemann_selftest_kernel");
}
