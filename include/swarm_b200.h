/*
 * swarm_b200.h -- C ABI of the B200-native batched drone-swarm step engine.
 *
 * Drop-in boundary for ONE hot path of nusRying/Multi-Agent-RL-for-Autonomous-Drone-Swarms:
 * the environment reset/step of the kinematic envs
 *     src/swarm_marl/envs/drone_swarm_env.py   DroneSwarmEnv   (reset :65-90, step :92-174)
 *     src/swarm_marl/envs/single_drone_env.py  SingleDroneEnv  (reset :53-71, step :73-111)
 *     src/swarm_marl/envs/common.py            DroneEnvConfig  (:7-32)
 * stepped for E independent env instances per call on one GPU.  The reference has no FFI (it
 * is pure Python); these entry points are what a ctypes binding of that path binds -- see
 * INTEGRATION.md for the reference-side stub.  No torch / C++ types cross this boundary:
 * plain pointers, sizes and PODs only.
 *
 * Ownership: every device buffer in SwarmBuffers is allocated and freed by the caller (the
 * Python host keeps them as torch CUDA tensors).  The library allocates device memory only in
 * swarm_create (a small PCG64 jump table + pinned/device staging for the host-buffer path)
 * and frees it in swarm_destroy.  Errors: 0 on success, negative SWARM_E_* otherwise; the
 * message is kept per thread in swarm_last_error().  No exceptions, no exit/abort.
 * Threading: all work is enqueued on the caller's stream without host synchronisation
 * (except swarm_step_host, which is synchronous by contract).  A handle is not thread-safe;
 * distinct handles are independent.
 */
#ifndef SWARM_B200_H
#define SWARM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SWARM_ABI_VERSION 5

enum {
    SWARM_OK = 0,
    SWARM_E_INVALID = -1,     /* bad argument / config value */
    SWARM_E_CUDA = -2,        /* a CUDA runtime call failed (see swarm_last_error) */
    SWARM_E_UNSUPPORTED = -3, /* config outside what the kernels are instantiated for */
    SWARM_E_NULL = -4         /* required pointer is NULL */
};

enum {
    SWARM_KIND_SINGLE = 0, /* SingleDroneEnv  (single_drone_env.py:12) */
    SWARM_KIND_SWARM = 1,  /* DroneSwarmEnv   (drone_swarm_env.py:17) */
    SWARM_KIND_PHYSICS = 2 /* DronePhysicsEnv (drone_physics_env.py:22) as a point mass: its env contract and the
                              force / drag / gravity / speed-clamp model around PyBullet's solver; any num_drones up
                              to SWARM_MAX_DRONES, domain randomisation allowed on top (round 2).  PyBullet is not
                              vendored: parity unpinned */
};

#define SWARM_MAX_DRONES 128    /* np.mean's pairwise sum is restated exactly for n <= 128 */
#define SWARM_MAX_NEIGHBOR_K 8
#define SWARM_MAX_SENSED 8

/* Mirror of DroneEnvConfig (envs/common.py:7-25) + num_drones (drone_swarm_env.py:32).
 * Scalars are doubles because the reference holds them as Python floats; the library derives
 * the float32 constants and thresholds the reference's numpy expressions effectively use. */
typedef struct SwarmConfig {
    int32_t abi_version; /* must be SWARM_ABI_VERSION */
    int32_t env_kind;    /* SWARM_KIND_* */
    int32_t num_envs;    /* E: independent env instances on this GPU */
    int32_t num_drones;  /* N (1 for SWARM_KIND_SINGLE) */
    int32_t num_obstacles;    /* M */
    int32_t sensed_obstacles; /* S */
    int32_t neighbor_k;       /* K (ignored for SWARM_KIND_SINGLE) */
    int32_t max_steps;
    int32_t norm_mode;  /* 0: np.linalg.norm(vec3) via BLAS sdot, double accumulate (x86-64 OpenBLAS);
                           1: sequential float32 accumulate */
    int32_t device;     /* CUDA device ordinal, -1 = current */
    double world_size, dt, max_speed, max_accel, collision_radius, goal_radius;
    double obstacle_radius, desired_spacing;
    double reward_progress_scale, reward_goal, reward_collision, reward_formation_scale;
    /* ---- domain randomisation (reference configs/domain_randomization_v1.yaml:9-60).  No reference
     * code consumes that file, so the semantics are this library's and parity is
     * pinned only against oracle/swarm_oracle.c; with dr_enabled == 0 nothing below is read and the
     * step is bit-identical to the reference. */
    int32_t dr_enabled;
    int32_t dr_reserved;
    uint64_t dr_seed;         /* key of the counter-based (Philox4x32-10) DR streams */
    int64_t env_index_base;   /* global index of this handle's env 0 (streams do not depend on sharding) */
    double dr_mass_scale[2], dr_max_accel_scale[2], dr_max_speed_scale[2], dr_dt_scale[2];
    double dr_obstacle_radius_scale[2], dr_world_size_scale[2];   /* uniform [min, max], per env per episode */
    double dr_thrust_noise_std;             /* multiplicative Gaussian on the clipped action, per drone per step */
    double dr_position_noise_std, dr_velocity_noise_std, dr_obstacle_distance_noise_std;  /* additive, on the obs */
    /* actuation.control_delay_steps: per episode a delay d is drawn from {values[k] with probability probs[k]};
     * the command applied at step t is the one submitted at step t - d (zero while the episode is younger) */
    int32_t dr_delay_count;       /* 0 = no control delay; <= 4 */
    int32_t dr_delay_reserved;
    int32_t dr_delay_values[4];   /* each in [0, 8] */
    double dr_delay_probs[4];
} SwarmConfig;

/* Element counts of every buffer, as a function of the config (swarm_query_sizes). */
typedef struct SwarmSizes {
    int64_t obs_dim;       /* D = 9 + 4K + 4S (swarm) | 9 + 4S (single) */
    int64_t state_dim;     /* 6N + 3 : length of one global_state row */
    int64_t pos4;          /* E*N*4  float32 */
    int64_t vel4;          /* E*N*4  float32 */
    int64_t goal4;         /* E*4    float32 */
    int64_t obst4;         /* E*M*4  float32 (>= 4 even when M == 0) */
    int64_t step_count;    /* E      int32 */
    int64_t rng;           /* E*4    uint64 */
    int64_t ep_return;     /* E      float32 */
    int64_t actions;       /* E*N*3  float32 */
    int64_t obs;           /* E*N*D  float32 */
    int64_t per_agent;     /* E*N    (reward, dist, flags) */
    int64_t per_env;       /* E      (all_terminated, all_truncated, episode outputs) */
    int64_t global_state;  /* E*(6N+3) float32 */
    int64_t stats;         /* SWARM_STATS_WORDS uint64 */
    int64_t dr_params;     /* E*8 float32 */
    int64_t act_hist;      /* E*H*N*3 float32, H = max(dr_delay_values) (0 when no control delay) */
} SwarmSizes;

/* Device pointers.  State is structure-of-arrays with one float4 per drone so every state
 * access is a coalesced 16-byte transaction:
 *   pos4[e][i] = (x, y, z, alive)   alive = 1.0f while drone_i is in env.agents, else 0.0f
 *   vel4[e][i] = (vx, vy, vz, 0)
 *   goal4[e]   = (gx, gy, gz, 0)        obst4[e][m] = (ox, oy, oz, 0)
 * Outputs marked "nullable" are skipped when the pointer is NULL. */
typedef struct SwarmBuffers {
    /* ---- state (read + written by reset/step) ---- */
    float *pos4;         /* [E][N][4] */
    float *vel4;         /* [E][N][4] */
    float *goal4;        /* [E][4] */
    float *obst4;        /* [E][M][4] */
    int32_t *step_count; /* [E]    DroneSwarmEnv.step_count */
    uint64_t *rng;       /* [E][4] numpy PCG64 {state_hi, state_lo, inc_hi, inc_lo} of env.rng */
    float *ep_return;    /* [E]    running sum of all agents' rewards in the current episode */
    /* ---- outputs of reset / observe / step ---- */
    float *obs;          /* [E][N][D]  observation rows, every drone (see obs_valid) */
    float *reward;       /* [E][N]     float32(reward); 0 for drones not stepped */
    double *reward64;    /* [E][N]     nullable: the reference's Python-float reward, bit-exact */
    float *dist;         /* [E][N]     info["distance_to_goal"] */
    uint8_t *terminated; /* [E][N] */
    uint8_t *truncated;  /* [E][N] */
    uint8_t *reached;    /* [E][N]     drone within goal_radius on this step */
    uint8_t *collision;  /* [E][N]     drone collided on this step */
    uint8_t *obs_valid;  /* [E][N]     1 where the reference puts drone_i into the obs/info dicts
                                       (after an auto-reset: 1 for every drone, obs = reset obs) */
    uint8_t *all_terminated; /* [E]    terminated["__all__"] */
    uint8_t *all_truncated;  /* [E]    truncated["__all__"] */
    float *global_state; /* [E][6N+3]  nullable: info["global_state"] = [pos.ravel|vel.ravel|goal] */
    float *episode_return;   /* [E]    nullable: return of the episode that ended on this step */
    int32_t *episode_length; /* [E]    nullable: its length in steps (0 when no episode ended) */
    uint64_t *stats;     /* [SWARM_STATS_WORDS] nullable: running counters, see SWARM_STAT_* */
    float *dr_params;    /* [E][8] state, required when dr_enabled: per-env per-episode constants
                            {max_accel, max_speed, dt, bound, obstacle threshold, episode key, world, control delay} */
    float *act_hist;     /* [E][H][N][3] state, required when a control delay is configured: command ring */
} SwarmBuffers;

/* stats block (uint64 words; SWARM_STAT_RETURN_SUM holds a double's bit pattern) */
enum {
    SWARM_STAT_EPISODES = 0,   /* episodes ended */
    SWARM_STAT_SUCCESS = 1,    /* ... with every drone at the goal and no collision */
    SWARM_STAT_COLLISION = 2,  /* ... by a collision */
    SWARM_STAT_TIMEOUT = 3,    /* ... by the time limit */
    SWARM_STAT_LENGTH_SUM = 4, /* sum of their lengths */
    SWARM_STAT_RETURN_SUM = 5, /* sum of their returns (double bits) */
    SWARM_STAT_AGENT_STEPS = 6,/* actions applied (one per active drone per step) */
    SWARM_STAT_ENV_STEPS = 7,  /* env instances stepped */
    SWARM_STAT_NAN_ACTIONS = 8,/* applied actions with a NaN component (np.clip lets NaN through, drone_swarm_env.py:105:
                                  the state of that env is NaN from then on -- this counter is the guard) */
    SWARM_STATS_WORDS = 9
};

/* Host-side output pointers for swarm_step_host (any may be NULL = do not copy back). */
typedef struct SwarmHostOut {
    float *obs; float *reward; double *reward64; float *dist;
    uint8_t *terminated; uint8_t *truncated; uint8_t *reached; uint8_t *collision; uint8_t *obs_valid;
    uint8_t *all_terminated; uint8_t *all_truncated;
    float *global_state;
    /* block mode (block_bytes > 0; the fields above are then ignored): the caller keeps every output buffer inside
     * ONE device allocation [block_dev, block_dev + block_bytes) and wants it mirrored byte for byte at block_host
     * -- one device->host copy per step instead of one per field (small batches: the E = 1 facade envs) */
    void *block_host; const void *block_dev; int64_t block_bytes;
    /* ABI 5: the five per-agent flag arrays as ONE byte per agent, packed on the device right behind the step
     * (SWARM_FLAG_* bits) -- a host consumer that wants every flag reads 1 byte per agent over the link instead
     * of 5.  Independent of the unpacked fields above (either, both or neither may be requested). */
    uint8_t *flags;      /* [E][N] */
} SwarmHostOut;

/* bits of SwarmHostOut.flags */
enum {
    SWARM_FLAG_TERMINATED = 1, SWARM_FLAG_TRUNCATED = 2, SWARM_FLAG_REACHED = 4, SWARM_FLAG_COLLISION = 8,
    SWARM_FLAG_OBS_VALID = 16
};

typedef struct SwarmHandle SwarmHandle;

int swarm_abi_version(void);
const char *swarm_last_error(void);

/* Replaces DroneSwarmEnv.__init__ / SingleDroneEnv.__init__ (drone_swarm_env.py:28-63,
 * single_drone_env.py:28-51) for E instances: validates the config, derives the float32
 * constants, builds the PCG64 jump table, picks the kernel instantiation. */
int swarm_create(const SwarmConfig *cfg, SwarmHandle **out);
int swarm_destroy(SwarmHandle *h);
int swarm_query_sizes(const SwarmConfig *cfg, SwarmSizes *out);

/* np.random.default_rng(seed) for each env (drone_swarm_env.py:35, :66-67): SeedSequence ->
 * PCG64 state into bufs->rng.  seeds: device [E] uint64; env_mask: device [E] uint8 or NULL. */
int swarm_seed(SwarmHandle *h, const SwarmBuffers *bufs, const uint64_t *seeds, const uint8_t *env_mask,
               void *stream);

/* env.reset() (drone_swarm_env.py:65-90 / single_drone_env.py:53-71) for the masked envs:
 * draws positions -> goal -> obstacles from each env's PCG64 stream, zeroes velocities and
 * step_count, writes obs / dist / obs_valid / global_state.  env_mask NULL = all envs. */
int swarm_reset(SwarmHandle *h, const SwarmBuffers *bufs, const uint8_t *env_mask, void *stream);

/* Recompute obs / dist / global_state from the current state (after the caller wrote
 * pos4 / vel4 / goal4 / obst4 directly -- state injection for parity runs). */
int swarm_observe(SwarmHandle *h, const SwarmBuffers *bufs, void *stream);

/* env.step(action_dict) (drone_swarm_env.py:92-174 / single_drone_env.py:73-111) for every env.
 * actions: device [E][N][3] float32 (a missing dict key is a zero row, :104).  auto_reset != 0:
 * an env whose episode ends is reset inside the same call -- by a second launch enqueued right
 * behind the step launch on the same stream -- from its own PCG64 stream, like `env.reset()`
 * without a seed; reward / flags keep the terminal step's values while
 * obs / dist / obs_valid / global_state describe the new episode. */
int swarm_step(SwarmHandle *h, const SwarmBuffers *bufs, const float *actions, int auto_reset, void *stream);

/* n_steps consecutive env.step() calls with ONE host call: step t applies actions[t] (device [n_steps][E][N][3]
 * float32); launches are enqueued back to back on `stream` (what SURVEY 7 calls step_many: small batches are
 * bound by the host's per-call cost, not by the device).  After the call the outputs describe the LAST step;
 * episode statistics (stats, ep_return) accumulate over all of them.  No host state changes per launch, so the
 * call -- like swarm_step -- can be captured into a CUDA graph and replayed. */
int swarm_step_many(SwarmHandle *h, const SwarmBuffers *bufs, const float *actions, int n_steps, int auto_reset,
                    void *stream);

/* Same step through HOST buffers (the end-to-end path): copies actions host->device, steps,
 * copies the requested outputs device->host, chunked over the env axis on internal streams so
 * the copies overlap the kernel, and returns when the host buffers are valid.  Host pointers
 * should be pinned for full PCIe rate.  `stream`: the caller's stream -- the step starts behind the work already
 * enqueued there (event, no device-wide synchronisation). */
int swarm_step_host(SwarmHandle *h, const SwarmBuffers *bufs, const float *actions_host,
                    const SwarmHostOut *out_host, int auto_reset, void *stream);

/* Number of kernel launches this handle has enqueued so far (bench bookkeeping). */
int64_t swarm_launch_count(const SwarmHandle *h);

/* The 256-entry half-normal quantile table q[m] = Phi^-1(0.5 + (m + 0.5) / 512) the DR noise is drawn from:
 * a 9-bit field (sign bit 8, m = bits 0-7) of a Philox block maps to +-q[m] (host, for tests). */
int swarm_dr_quantile_table(float *out256);

#ifdef __cplusplus
}
#endif
#endif /* SWARM_B200_H */
