/*
 * doper_b200 — C ABI of the B200-native differentiable ball-rollout path.
 *
 * This is the drop-in boundary for the one hot path of VGGVRobotics/doper: the
 * doper/sim step + rollout and its reverse-mode backward.  Every entry point is
 * `extern "C"`, takes plain device pointers, sizes and a CUDA stream, returns an
 * int status (0 = ok), never aborts, never synchronises, allocates nothing and
 * keeps no global mutable state: the caller owns every buffer including the
 * workspaces whose sizes the *_bytes() queries return (one explicit exception:
 * doper_peer_alloc, because CUDA IPC needs memory from cudaMalloc).  All launches go to the
 * caller's stream; the library is re-entrant (one host thread per GPU is safe).
 *
 * Reference interfaces replaced (file:line in the reference tree):
 *   doper_scene_prepare        create_scene                     doper/sim/jax_geometry.py:302-305
 *   doper_closest_segment      find_closest_segment_to_point(s_batch)
 *                                                               doper/sim/jax_geometry.py:159-177
 *   doper_points_inside        if_points_inside_any_polygon     doper/sim/jax_geometry.py:97-116
 *   doper_rollout_fwd          run_sim / _run_sim (lax.scan of sim_step)
 *                                                               doper/sim/ball_sim.py:171-253,327
 *   doper_l1_loss              _compute_loss / reduce_loss      doper/sim/ball_sim.py:256-324
 *   doper_rollout_bwd          the VJP of the scan that jax.value_and_grad builds
 *                                                               doper/sim/ball_sim.py:334-338
 *   doper_value_and_grad       vmapped_grad_and_value           doper/sim/ball_sim.py:334-338
 *   doper_ray_segment_points   batch_line_ray_intersection_point doper/sim/jax_geometry.py:237-261
 *   doper_range_sensor         (Undirected/Simple)RangeSensor.get_observation
 *                                                               doper/agents/observations.py:29-110
 *   doper_agent_observations   RangeSensingAgent.get_observations
 *                                                               doper/agents/range_sensor_agent.py:17-41
 *
 * Data layout (all float = IEEE binary32, row-major, device memory):
 *   segments  [S][2][2] = (x0,y0,x1,y1) per segment, polygon-major, 3 per triangle;
 *             per-environment scenes are [E][S][2][2].
 *   points / x0 / v0 / xT / vT / cotangents   [B][2]
 *   trajectory (optional)  traj_x, traj_v, traj_a  [n_steps][B][2]
 *   hit log   int32 : bit 31 = the step collided; bits 30 / 29 = |v_y| < 1 / |v_x| < 1 at the start
 *             of the step (the friction clip's derivative: all the adjoint of a free-flight step
 *             needs); bits 0-28 = payload.  Step without a collision: closest segment index with
 *             flags bit 17, else 0.  Collided step: slot of the step's collision record, or, with
 *             bit 28 set, the segment index itself (no record could be stored; the backward then
 *             replays the step from its checkpoint).
 *             [n_steps][B] or [B][n_steps] (see doper_rollout_log_ball_major)
 *   collision records  32 B each, appended by the forward in arrival order: float (x,y,vx,vy) at the
 *             START of the collided step, then int32 (ball low word, step, segment index, ball high
 *             word).  The int32 at offset 0 of the workspace counts the slots handed out; a slot that no log
 *             word refers to (the unused rest of a slab a tail worker reserved) holds stale bytes.
 *   checkpoints [ceil(n_steps/ckpt_every)][B][4] = (x,y,vx,vy) at the START of step k*ckpt_every
 *   workspace (doper_rollout_workspace_bytes): [256 B header | hand-over flags u32[B] | checkpoints | hit log |
 *             records | hand-over entries 32 B x B | 256 B], each part 256-B aligned; written by
 *             doper_rollout_fwd only.  Header: int32 record count at offset 0, finished-block counter at 16,
 *             hand-over queue counters at 128 (entries pushed) and 132 (entries claimed by tail workers).
 *   backward scratch (doper_rollout_bwd_scratch_bytes): one 4x4 binary64 Jacobian per record
 *             capacity; written by doper_rollout_bwd (so that the forward's workspace stays
 *             read-only for the backward: XLA input buffers are immutable, and two backward calls
 *             may share one forward on different streams with different scratch buffers).
 *   consts    host float[5] = {radius, mass, rolling_friction_coefficient,
 *             walls_elasticity, attractor_strength}  (reference constants dict,
 *             doper/sim/ball_sim.py:104,119,127-129,195-200); g = 9.8f is the literal
 *             default of compute_rolling_friction_force (ball_sim.py:29).
 *   attractor / target  host float[2].
 */
#ifndef DOPER_B200_H
#define DOPER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* doper_stream_t; /* cudaStream_t */

enum {
    DOPER_OK = 0,
    DOPER_ERR_BAD_SHAPE = 1,   /* negative / zero / inconsistent sizes */
    DOPER_ERR_NULL_PTR = 2,    /* a required pointer is NULL */
    DOPER_ERR_MISALIGNED = 3,  /* pointer not aligned as documented */
    DOPER_ERR_SMEM_LIMIT = 4,  /* scene does not fit the kernel's shared-memory plan (pinned full-sweep rollout kernels
                                  only: the default rollout plans, the queries and the sensor read a large scene in place) */
    DOPER_ERR_LAUNCH = 5,      /* cudaPeekAtLastError() != success after launch */
    DOPER_ERR_WORKSPACE = 6,   /* workspace too small */
    DOPER_ERR_UNSUPPORTED = 7  /* e.g. S >= 2^31, triangle count mismatch */
};

#define DOPER_ABI_VERSION 3

int doper_abi_version(void);

/* The floating-point contraction mask the kernels were compiled with (include/doper_fp_policy.h): which
 * mul+add pairs of the reference's arithmetic are single-rounding FMAs.  Default 1 (the coordinate update). */
int doper_fp_policy(void);
const char* doper_error_string(int code);

/* ---- scene ---------------------------------------------------------------------------- */

/* Bytes of the prepared-scene workspace for n_scenes scenes of S segments each (256-B aligned
 * device buffer owned by the caller). */
size_t doper_scene_workspace_bytes(int64_t n_scenes, int32_t S);

/* Pre-digest raw segments into the kernels' staging layout: per segment (s0x,s0y,svx,svy) and
 * 1/|sv|^2, padded to a multiple of 32 with far-away sentinels, the raw end points (read by the binary64
 * backward), plus a per-scene header (max |coordinate|, degenerate-segment flag).  Bit-identical to
 * computing sv/L per pair as the reference does (jax_geometry.py:132-133): both are ball-independent.
 * n_scenes = 1 for a shared scene (reference semantics, ball_sim.py:332) or E for per-environment scenes.
 * max_radius > 0 (shared scenes): also bake the broad phase — a uniform grid over the scene whose cells
 * hold every segment a ball of radius <= max_radius can touch from anywhere in the cell — which the
 * default forward kernel of shared scenes reads instead of scanning the scene.  A rollout whose radius
 * exceeds max_radius (or max_radius <= 0) still returns the same results, through the reference
 * arithmetic over every segment (slow): prepare the scene with the radius you roll out with. */
int doper_scene_prepare(doper_stream_t stream, int64_t n_scenes, int32_t S, const float* segments,
                        float max_radius, void* scene_ws);

/* ---- geometry queries ----------------------------------------------------------------- */

/* idx[i] = first index attaining min_j dist(points[i], segment j); dist[i] that distance
 * (NaN distances count as minimal, like argmin). seg_out (nullable) [N][2][2] = segments[idx]. */
int doper_closest_segment(doper_stream_t stream, int64_t N, int32_t S, const float* points,
                          const float* segments, const void* scene_ws, int32_t* idx, float* dist,
                          float* seg_out);

/* flag[i] = 1 iff points[i] is strictly inside any of the S/3 counter-clockwise triangles. */
int doper_points_inside(doper_stream_t stream, int64_t N, int32_t S, const float* points,
                        const float* segments, uint8_t* flag);

/* Per-environment forms: point i is tested against scene i (segments [N][S][2][2], scene_ws prepared with
 * n_scenes = N); the scenes are read in place.  Same results as N calls with one point each. */
int doper_closest_segment_env(doper_stream_t stream, int64_t N, int32_t S, const float* points,
                              const float* segments, const void* scene_ws, int32_t* idx, float* dist,
                              float* seg_out);
int doper_points_inside_env(doper_stream_t stream, int64_t N, int32_t S, const float* points,
                            const float* segments, uint8_t* flag);

/* ---- range sensor (observation path) ------------------------------- */

/* batch_line_ray_intersection_point (doper/sim/jax_geometry.py:237-261):
 * points[r][s] = intersection of ray r (origin, direction; the direction need not be normalised)
 * with segment s, or (inf, inf) when they do not meet.  origins / directions [R][2], points [R][S][2]. */
int doper_ray_segment_points(doper_stream_t stream, int64_t R, int32_t S, const float* origins,
                             const float* directions, const float* segments, float* points);

/* SimpleRangeSensor / UndirectedRangeSensor.get_observation (doper/agents/observations.py:29-110) for N
 * sensor positions at once, all with the same R ray directions: ranges[i][r] = distance from
 * positions[i] to the closest intersection along ray r, clamped to distance_range; points (nullable)
 * [N][R][2] that intersection (inf beyond distance_range); idx (nullable) [N][R] its segment
 * (first index on ties, 0 when nothing is hit). */
int doper_range_sensor(doper_stream_t stream, int64_t N, int32_t R, int32_t S, const float* positions,
                       const float* ray_directions, const void* scene_ws, float distance_range, float* ranges,
                       float* points, int32_t* idx);

/* Per-environment form: position i looks at scene i (scene_ws prepared with n_scenes = N). */
int doper_range_sensor_env(doper_stream_t stream, int64_t N, int32_t R, int32_t S, const float* positions,
                           const float* ray_directions, const void* scene_ws, float distance_range, float* ranges,
                           float* points, int32_t* idx);

/* float64 positions (doper/agents/observations.py:50,67: numpy keeps a float64 position as it is, so the
 * reference's ranges `norm(points - position)` are float64, while JAX rounds the ray origins to float32 for
 * the intersections, jax_geometry.py:237-261): intersections from the rounded origins, ranges and the
 * (range, first index) minimum in binary64, float64 ranges out.  per_env = 0 / 1 as above. */
int doper_range_sensor_f64(doper_stream_t stream, int64_t N, int32_t R, int32_t S, int32_t per_env,
                           const double* positions, const float* ray_directions, const void* scene_ws,
                           float distance_range, double* ranges, float* points, int32_t* idx);

/* RangeSensingAgent.get_observations (doper/agents/range_sensor_agent.py:17-41), fused:
 * observations[i] = [ |x_i - target|, (x_i - target) / |x_i - target| (2), x_i (2), v_i (2),
 * ranges[i][0..R) / distance_range ], float32 [N][7 + R]; the first three entries are evaluated in
 * float64 like the reference's numpy code (its target is a float64 array) and cast once. */
int doper_agent_observations(doper_stream_t stream, int64_t N, int32_t R, int32_t S, const float* positions,
                             const float* velocities, const float* ray_directions, const void* scene_ws,
                             float distance_range, const double target[2], float* observations);
int doper_agent_observations_env(doper_stream_t stream, int64_t N, int32_t R, int32_t S, const float* positions,
                                 const float* velocities, const float* ray_directions, const void* scene_ws,
                                 float distance_range, const double target[2], float* observations);

/* ---- rollout -------------------------------------------------------------------------- */

typedef struct {
    int64_t B;           /* balls (= environments when per_env) */
    int32_t S;           /* segments per scene (<= 65,504: DOPER_ERR_UNSUPPORTED beyond) */
    int32_t per_env;     /* 0: one shared scene; 1: scene b belongs to ball b */
    int32_t n_steps;     /* rollout length */
    int32_t ckpt_every;  /* K: checkpoint period in steps (>= 1); 0 = library default */
    float dt;            /* f32(sim_time / n_steps), rounded on the host (ball_sim.py:237) */
    float consts[5];     /* radius, mass, friction, elasticity, attractor_strength */
    float attractor[2];
    int32_t lanes_per_ball; /* 0 = auto; else 1,2,4,8,16,32 lanes cooperating on one ball's sweep */
    int32_t flags;          /* bit 0: force the exact (unfiltered) sweep;
                               bit 1: the sequential backward: fp32 in the reference's operation
                                      order, one thread per ball (bit-identical to the oracle's
                                      adjoint).  Default: the binary64 chunk-parallel backward (the
                                      adjoint of the real-arithmetic step on the fp32 trajectory,
                                      rounded once to fp32 at the end);
                               bit 2: force the shared-memory lane-split forward kernel over the
                                      register-resident one (only with lanes_per_ball != 0);
                               bit 3: disable the time-parallel (warp per ball) full-sweep forward kernel;
                               bit 15 / bit 16: forbid / force the broad-phase (candidate list)
                                      forward kernels; bit 18: with bit 16 and lanes_per_ball >= 2, the
                                      time-parallel (lane = step) variant instead of the lane-split one;
                               bit 21: with lanes_per_ball = 1, 2, 4, 8 or 16: the baked kernel with that many
                                      lanes (speculated steps) per ball instead of the batch-size rule;
                               bit 20: keep the per-ball candidate-list kernels for a shared scene
                                      (default there: the baked per-cell lists of doper_scene_prepare);
                               bit 22: forbid the pipelined forward kernel (producer / consumer warp pairs over
                                      the baked lists: the default of shared scenes up to 6,144 balls);
                               bit 24: doper_value_and_grad keeps forward, loss and backward as separate launches
                                      (default up to 1,024 balls on the pipelined forward: ONE launch, each block
                                      pulls the gradient of its own balls back as soon as they have finished);
                               bit 25: baked forward kernels (not the pipelined one): keep a ball in sliding contact
                                      in its warp (default: after 8 collided steps in a row it is handed to a
                                      tail-worker warp that runs it alone — its serial chain, lengthened by its
                                      warp mates' divergent work, otherwise sets the duration of the whole launch);
                               bit 23: with lanes_per_ball = 8, 16 or 32: the pipelined kernel with that window
                                      length (speculated steps per producer / consumer hand-over);
                               bit 19: no collision records (the backward replays every collided
                                      step from its checkpoint: slower, same results);
                               bit 17: hit log holds the exact closest index at EVERY step (the
                                      broad-phase kernel otherwise logs it at collisions only and 0
                                      elsewhere: the index is consumed at collisions only); the
                                      candidate lists then hold the argmin of every point of their
                                      disc instead of just the segments the ball can touch (slower).
                               None of these changes x_T, v_T, trajectories, hit flags or collision
                               indices; bit 1 selects the rounding of the gradient (see above). */
    int64_t hit_capacity;   /* collision records the workspace can hold; 0 = library default
                               (max(65,536, B n_steps / 64), at most B n_steps) */
} doper_rollout_desc;

/* The checkpoint period the library will use for this rollout (ckpt_every if > 0, else 32). */
int32_t doper_rollout_ckpt_every(const doper_rollout_desc* d);

/* Layout of the hit log inside the workspace: 0 = [n_steps][B] (step-major), 1 = [B][n_steps]
 * (ball-major, written by the time-parallel forward kernel). Only tools that decode the
 * workspace need this; doper_rollout_bwd derives it from the same descriptor. */
int32_t doper_rollout_log_ball_major(const doper_rollout_desc* d);

/* Name of the kernel the library will launch for this rollout (backward = 0: doper_rollout_fwd, 1: doper_rollout_bwd,
 * 2: the backward inside doper_value_and_grad, which may be fused into the forward launch; 3: the second kernel
 * doper_rollout_fwd launches behind the main one — the hand-over drain of the one-thread-per-ball plan — or "");
 * a static string, for benchmark / profile reports. */
const char* doper_rollout_plan_name(const doper_rollout_desc* d, int32_t backward);

/* Bytes of the forward's workspace (header + checkpoints + hit log + collision records). */
size_t doper_rollout_workspace_bytes(const doper_rollout_desc* d);

/* Bytes of the backward's scratch buffer (256-B aligned device memory, contents undefined on entry). */
size_t doper_rollout_bwd_scratch_bytes(const doper_rollout_desc* d);

/* Forward rollout.  traj_* and ws may be NULL (no trajectory / no backward wanted). */
int doper_rollout_fwd(doper_stream_t stream, const doper_rollout_desc* d, const void* scene_ws,
                      const float* x0, const float* v0, float* xT, float* vT, float* traj_x,
                      float* traj_v, float* traj_a, void* ws);

/* Backward: cotangents (xT_bar, vT_bar) -> (x0_bar, v0_bar).  vT_bar / x0_bar may be NULL.
 * ws is the forward's workspace (read only); bwd_scratch is written (it may be NULL with flags
 * bit 1, bit 17 or bit 19, where doper_rollout_bwd_scratch_bytes has nothing to hold). */
int doper_rollout_bwd(doper_stream_t stream, const doper_rollout_desc* d, const void* scene_ws,
                      const void* ws, void* bwd_scratch, const float* xT_bar, const float* vT_bar,
                      float* x0_bar, float* v0_bar);

/* loss_per_ball[b] = |xT[b].x - t.x| + |xT[b].y - t.y|; *loss_sum = sum_b (deterministic order);
 * xT_bar[b] = sign(xT[b] - t) (the seed of the backward).  loss_per_ball / xT_bar nullable. */
int doper_l1_loss(doper_stream_t stream, int64_t B, const float* xT, const float target[2],
                  float* loss_per_ball, float* loss_sum, float* xT_bar);

/* vmapped_grad_and_value in one call: fwd -> loss -> bwd on `stream` — as ONE kernel launch for small batches on the
 * pipelined forward plan (shared scene in shared memory, <= 1,024 balls, default backward, collision records on; flags
 * bit 24 forces the separate launches), else a forward, a loss and one or two backward launches.  Same results either way
 * (the loss total is summed in the same order).  scratch: device float[2*B] (holds xT_bar / per-ball loss terms).
 * x0_bar and loss_per_ball nullable. */
int doper_value_and_grad(doper_stream_t stream, const doper_rollout_desc* d, const void* scene_ws,
                         const float* x0, const float* v0, const float target[2], void* ws, void* bwd_scratch,
                         float* scratch, float* loss_sum, float* loss_per_ball, float* xT,
                         float* vT, float* v0_bar, float* x0_bar);

/* ---- multi-GPU exchange --------------------------------------- */

/* One-shot all-reduce(sum) of n floats over NVLink peer memory, for the one exchange of the data-parallel
 * trainer (controller gradients + loss, 26 KB: latency bound).  Every rank (one process per GPU) calls
 * doper_peer_alloc once (the only allocation the library ever makes: an explicit, caller-owned buffer
 * that must be cudaMalloc'ed for CUDA IPC), exchanges the 64-byte handles with its peers by any host
 * channel (the Python binding uses torch.distributed), and opens theirs with doper_peer_open.  Per step
 * every rank calls doper_allreduce_oneshot ONCE (the same number of calls on every rank — the step
 * counter lives in the buffer, so the arguments never change and the call can sit in a CUDA graph):
 * out[i] = sum over ranks of in[i], added in rank order (the same bits everywhere).  `in` and `out` are
 * ordinary device arrays of n floats (they may alias); `buffers` is a HOST array of `world` device
 * pointers indexed by rank (own buffer at [rank]).  One kernel: it pushes `in` into every peer with the
 * step number inside each 16-byte store and polls its own buffer for the peers' words (no fence, no
 * separate flag: one NVLink crossing).  A peer that never arrives makes the kernel give up after
 * timeout_seconds (<= 0: one minute), set *status (device int, nullable) to 1 and fill `out` with NaN instead
 * of spinning for ever; the caller must check *status at a synchronisation point before trusting parameters
 * updated from `out` (the Python binding does: distributed.PeerAllReducer.check()).
 * Memory-model assumption (csrc/peer.cuh): the {value, step} halves of a 16-byte volatile vector store to
 * peer memory are observed untorn (8-byte granularity) by the peer's 16-byte volatile loads.  True on NVLink 5 /
 * NVSwitch (verified against NCCL by bench.py's multi-GPU self-check on every run), not promised by PTX. */
size_t doper_peer_buffer_bytes(int32_t n);
int doper_peer_alloc(int32_t n, void** buffer, unsigned char handle[64]);
int doper_peer_open(const unsigned char handle[64], void** buffer);
int doper_peer_close(void* buffer);
int doper_peer_free(void* buffer);
int doper_allreduce_oneshot(doper_stream_t stream, int32_t rank, int32_t world, void* const* buffers, int32_t n,
                            const float* in, float* out, int32_t* status, double timeout_seconds);

/* ---- trainer glue around the rollout ------------------------------------ */

/* The reference's controller (doper/networks/controllers.py:12-18): Linear(d_in, h1) - ReLU - Linear(h1, h2) - ReLU -
 * Linear(h2, d_out), dims = {d_in, h1, h2, d_out} (shipped: 25, 50, 100, 2), parameters as ONE flat float vector in
 * torch order W1 [h1][d_in], b1, W2 [h2][h1], b2, W3 [d_out][h2], b3.
 *   doper_controller_fwd   impulse [B][d_out] = controller(observations [B][d_in]) (nullable), and
 *                          v0 [B][2] = velocity_init + impulse[:, :2] / mass (nullable)      ball_trainer.py:57-64
 *   doper_controller_bwd   grad [n_params] += d(sum_b v0_bar[b] . v0[b]) / d params = impulse.backward(dL/dv0)
 *                          (ball_trainer.py:70-71); deterministic (block partials in `scratch`,
 *                          doper_controller_scratch_bytes, added in block order); `grad` can be the buffer that
 *                          doper_allreduce_oneshot exchanges
 *   doper_adam_clip_step   torch.nn.utils.clip_grad_norm_(max_norm) + torch.optim.Adam.step() (no weight decay / amsgrad)
 *                          on the flat vectors, then grad <- 0; `step` is a device float counter (starts at 0);
 *                          total_norm (nullable) receives the gradient norm before clipping   ball_trainer.py:74-76 */
size_t doper_controller_scratch_bytes(const int32_t dims[4], int64_t B);
int doper_controller_fwd(doper_stream_t stream, const int32_t dims[4], int64_t B, const float* params, const float* observations,
                         const float* velocity_init, float mass, float* impulse, float* v0);
int doper_controller_bwd(doper_stream_t stream, const int32_t dims[4], int64_t B, const float* params, const float* observations,
                         const float* v0_bar, float mass, void* scratch, float* grad);
int doper_adam_clip_step(doper_stream_t stream, int32_t n, float* params, float* grad, float* exp_avg, float* exp_avg_sq, float* step,
                         float max_norm, float lr, float beta1, float beta2, float eps, float* total_norm);

/* ---- measurement helper --------------------------------------------------------------- */

/* Register-only dependent FADD/FMUL chains (no FMA): every thread executes `iters` x 16
 * independent-chain ops; out[thread] keeps the result alive.  Used by bench.py to measure the
 * non-FMA FP32 lane-op rate that bounds a bit-exact sweep. */
int doper_fp32_probe(doper_stream_t stream, int32_t blocks, int32_t threads, int32_t iters,
                     int32_t use_fma, float* out);

#ifdef __cplusplus
}
#endif
#endif /* DOPER_B200_H */
