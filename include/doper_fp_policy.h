/*
 * doper_fp_policy.h — the ONE place that says which multiply-add pairs of the reference's fp32
 * arithmetic are evaluated as a single-rounding fused multiply-add.
 *
 * Why this exists.  The reference (doper/sim/ball_sim.py, doper/sim/jax_geometry.py) is a JAX program:
 * what "a*b + c" rounds to is decided by XLA:CPU / LLVM (jax==0.1.68, jaxlib==0.1.47, not installable
 * here), not by the Python source.  The reference's only recorded output (notebook cell 7,
 * tests/golden/notebook_cell7.json) proves that its compiler DID contract the coordinate update
 * x + v'*dt (ball_sim.py:76) and did NOT contract the velocity update v + a*dt (ball_sim.py:75) nor
 * ks*p + f (ball_sim.py:200); nothing recorded pins the collision path (jax_geometry.py:132-136,
 * ball_sim.py:97-132).  The default policy below contracts exactly what is pinned and evaluates the rest
 * as written.  tests/test_fp_policy_cpu.py measures how many collision decisions of the C2 workload
 * depend on the unpinned choices; a future image with the reference's JAX settles them, and then the
 * flip is one number here (or -DDOPER_FP_POLICY=<mask> on both command lines).
 *
 * Included by BOTH sides, so that they cannot drift apart:
 *   doper_b200/csrc/physics.cuh, sweep.cuh, queries.cuh   (the product kernels, nvcc)
 *   oracle/ball_sim_ref.c + ball_sim_step.inc             (the CPU checker, gcc -ffp-contract=off)
 * oracle/ball_sim_np.py mirrors the DEFAULT mask only (asserted equal to the C oracle in tests/).
 *
 * Bits (reference lines each one touches):
 */
#ifndef DOPER_FP_POLICY_H
#define DOPER_FP_POLICY_H

#define DOPER_FP_POS 1      /* x + v*dt            ball_sim.py:76 (get_new_state), :110, :132        PINNED on  */
#define DOPER_FP_VEL 2      /* v + a*dt            ball_sim.py:75, :109, :131                        PINNED off */
#define DOPER_FP_LERP 4     /* s0 + t*sv           jax_geometry.py:136                               unpinned   */
#define DOPER_FP_DOT 8      /* a0*b0 + a1*b1 -> fma(a0, b0, a1*b1): every 2-vector dot product / squared norm:
                               jax_geometry.py:133-134,152 ; ball_sim.py:101-103,113-116            unpinned   */
#define DOPER_FP_DOT_REV 16 /* with DOT: the other product is the fused one, fma(a1, b1, a0*b0)      unpinned   */
#define DOPER_FP_FORCE 32   /* ks*p + f            ball_sim.py:200 (PINNED off by the notebook) and
                               vp - e*vn           ball_sim.py:119 (unpinned)                                   */

#ifndef DOPER_FP_POLICY
#define DOPER_FP_POLICY DOPER_FP_POS
#endif

#if defined(__CUDACC__)
#define DOPER_FMAF(a, b, c) __fmaf_rn((a), (b), (c))
#else
#include <math.h>
#define DOPER_FMAF(a, b, c) fmaf((a), (b), (c))
#endif

/* Every macro below is for binary32 operands; both translation units are compiled without implicit
 * contraction (nvcc -fmad=false, gcc -ffp-contract=off), so the "as written" forms round twice. */
#if (DOPER_FP_POLICY) & DOPER_FP_POS
#define DOPER_POS_UPDATE(v, dt, x) DOPER_FMAF((v), (dt), (x))
#else
#define DOPER_POS_UPDATE(v, dt, x) ((x) + (v) * (dt))
#endif

#if (DOPER_FP_POLICY) & DOPER_FP_VEL
#define DOPER_VEL_UPDATE(a, dt, v) DOPER_FMAF((a), (dt), (v))
#else
#define DOPER_VEL_UPDATE(a, dt, v) ((v) + (a) * (dt))
#endif

#if (DOPER_FP_POLICY) & DOPER_FP_LERP
#define DOPER_LERP(s0, t, sv) DOPER_FMAF((t), (sv), (s0))
#else
#define DOPER_LERP(s0, t, sv) ((s0) + (t) * (sv))
#endif

#if ((DOPER_FP_POLICY) & DOPER_FP_DOT) && ((DOPER_FP_POLICY) & DOPER_FP_DOT_REV)
#define DOPER_DOT2(a0, b0, a1, b1) DOPER_FMAF((a1), (b1), (a0) * (b0))
#elif (DOPER_FP_POLICY) & DOPER_FP_DOT
#define DOPER_DOT2(a0, b0, a1, b1) DOPER_FMAF((a0), (b0), (a1) * (b1))
#else
#define DOPER_DOT2(a0, b0, a1, b1) ((a0) * (b0) + (a1) * (b1))
#endif

#if (DOPER_FP_POLICY) & DOPER_FP_FORCE
#define DOPER_FORCE_SUM(ks, p, f) DOPER_FMAF((ks), (p), (f))
#define DOPER_REFLECT(vp, e, vn) DOPER_FMAF(-(e), (vn), (vp))
#else
#define DOPER_FORCE_SUM(ks, p, f) ((ks) * (p) + (f))
#define DOPER_REFLECT(vp, e, vn) ((vp) - (e) * (vn))
#endif

#endif /* DOPER_FP_POLICY_H */
