// swarm_abi.cu -- the extern "C" surface declared in include/swarm_b200.h.
// Host logic only: config validation, derivation of the float32 constants the reference's numpy
// expressions effectively use, PCG64 jump table, launch geometry, the chunked host-buffer path.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "swarm_internal.h"

using namespace swarm;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            return fail(SWARM_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

#ifndef SWARM_HOST_CHUNKS
#define SWARM_HOST_CHUNKS 4
#endif
constexpr int kHostChunks = SWARM_HOST_CHUNKS;

}  // namespace

struct SwarmHandle {
    SwarmConfig cfg;
    DevParams base;          // everything but the buffer pointers / per-launch fields
    int device;
    int num_sms;
    int blocks_per_sm;
    size_t smem_bytes;
    bool rot_ok;             // the step launches run on swarm_step_rot_kernel (swarm_step_rot.cu)
    bool rot_fused;          // ... with the auto-reset inside the step launch (SWARM_B200_FUSED_RESET=0: second launch)
    int rot_blocks_per_sm;
    int multi_step_max_groups;   // swarm_step_many: batches up to this many env groups run all steps in ONE launch
    bool rotx_ok;            // N = 64 / 128: swarm_step_rotx_kernel (swarm_step_rotx.cu)
    int rotx_blocks_per_sm;        // step launch
    int rotx_reset_blocks_per_sm;  // auto-reset launch (lighter kernel, more CTAs per SM)
    int rotx_min_envs;       // launches with fewer envs stay on the general kernel (SWARM_B200_ROTX_MIN_ENVS)
    JumpEntry* jump_dev;
    uint8_t* reset_mask_dev;  // [E]
    unsigned* reset_count_dev;  // [(kHostChunks + 1) * 2] per launch slot: the counters of the two lists
    unsigned* reset_epoch_dev;  // [kHostChunks + 1] per launch slot: which list is current (device-side parity)
    int* reset_list_dev;        // [2][number of groups + 1]
    int reset_list_stride;
    unsigned* work_counter_dev; // [(kHostChunks + 1) * 2] group queue of the rotation-pass step kernel, per launch slot
    float* qtable_dev;          // [512] signed quantile table (domain randomisation only)
    int64_t launches;
    // host-buffer path
    cudaStream_t chunk_stream[kHostChunks];
    cudaEvent_t chunk_done[kHostChunks];
    cudaEvent_t caller_ready;   // recorded on the caller's stream: the chunk streams start behind it
    float* actions_dev;      // [E][N][3] staging for swarm_step_host
    uint8_t* flags_dev;      // [E][N] packed flag bytes (SwarmHostOut.flags), allocated on first use
    bool host_path_ready;
};

namespace {

// SwarmHostOut.flags: terminated | truncated << 1 | reached << 2 | collision << 3 | obs_valid << 4, one byte per agent
// (host-buffer path only: runs on the chunk's stream between the step and the device->host copies; 4 agents per
//  thread when the chunk is 4-byte aligned)
__global__ void swarm_pack_flags_kernel(const uint8_t* __restrict__ term, const uint8_t* __restrict__ trunc,
                                        const uint8_t* __restrict__ reached, const uint8_t* __restrict__ col,
                                        const uint8_t* __restrict__ valid, uint8_t* __restrict__ out, long long a0,
                                        long long n) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (((a0 | n) & 3) == 0) {
        if (t * 4 >= n) return;
        const long long w = (a0 >> 2) + t;
        const unsigned m = 0x01010101u;
        const unsigned v = (reinterpret_cast<const unsigned*>(term)[w] & m) | ((reinterpret_cast<const unsigned*>(trunc)[w] & m) << 1) |
                           ((reinterpret_cast<const unsigned*>(reached)[w] & m) << 2) |
                           ((reinterpret_cast<const unsigned*>(col)[w] & m) << 3) |
                           ((reinterpret_cast<const unsigned*>(valid)[w] & m) << 4);
        reinterpret_cast<unsigned*>(out)[w] = v;
    } else {
        if (t >= n) return;
        const long long a = a0 + t;
        out[a] = (uint8_t)((term[a] & 1) | ((trunc[a] & 1) << 1) | ((reached[a] & 1) << 2) | ((col[a] & 1) << 3) |
                           ((valid[a] & 1) << 4));
    }
}

int obs_dim_of(const SwarmConfig& c) {
    return c.env_kind != SWARM_KIND_SINGLE ? 9 + 4 * c.neighbor_k + 4 * c.sensed_obstacles
                                          : 9 + 4 * c.sensed_obstacles;
}

// ring size of the command history: the largest configured control delay (0 = no delay anywhere)
int delay_hist_of(const SwarmConfig& c) {
    int h = 0;
    if (c.dr_enabled)
        for (int k = 0; k < c.dr_delay_count && k < 4; ++k) h = c.dr_delay_values[k] > h ? c.dr_delay_values[k] : h;
    return h;
}

int validate(const SwarmConfig* c) {
    if (!c) return fail(SWARM_E_NULL, "config is NULL");
    if (c->abi_version != SWARM_ABI_VERSION)
        return fail(SWARM_E_INVALID, "abi_version %d != %d", c->abi_version, SWARM_ABI_VERSION);
    if (c->env_kind != SWARM_KIND_SINGLE && c->env_kind != SWARM_KIND_SWARM && c->env_kind != SWARM_KIND_PHYSICS)
        return fail(SWARM_E_INVALID, "env_kind %d unknown", c->env_kind);
    if (c->num_envs < 1) return fail(SWARM_E_INVALID, "num_envs must be >= 1");
    if (c->num_drones < 1) return fail(SWARM_E_INVALID, "num_drones must be >= 1");
    if (c->env_kind == SWARM_KIND_SINGLE && c->num_drones != 1)
        return fail(SWARM_E_INVALID, "SWARM_KIND_SINGLE needs num_drones == 1");
    if (c->num_drones > SWARM_MAX_DRONES)
        return fail(SWARM_E_UNSUPPORTED, "num_drones %d > %d", c->num_drones, SWARM_MAX_DRONES);
    if (c->num_obstacles < 0 || c->num_obstacles > 64)
        return fail(SWARM_E_UNSUPPORTED, "num_obstacles %d outside [0, 64]", c->num_obstacles);
    if (c->sensed_obstacles < 0 || c->sensed_obstacles > SWARM_MAX_SENSED)
        return fail(SWARM_E_UNSUPPORTED, "sensed_obstacles %d outside [0, %d]", c->sensed_obstacles, SWARM_MAX_SENSED);
    if (c->neighbor_k < 0 || c->neighbor_k > SWARM_MAX_NEIGHBOR_K)
        return fail(SWARM_E_UNSUPPORTED, "neighbor_k %d outside [0, %d]", c->neighbor_k, SWARM_MAX_NEIGHBOR_K);
    if (c->norm_mode != 0 && c->norm_mode != 1) return fail(SWARM_E_INVALID, "norm_mode must be 0 or 1");
    if (c->dr_enabled) {
        if (c->norm_mode != 0) return fail(SWARM_E_UNSUPPORTED, "domain randomisation needs norm_mode 0");
        const double* rng[6] = {c->dr_mass_scale, c->dr_max_accel_scale, c->dr_max_speed_scale, c->dr_dt_scale,
                                c->dr_obstacle_radius_scale, c->dr_world_size_scale};
        for (int k = 0; k < 6; ++k)
            if (!(rng[k][0] > 0.0) || !(rng[k][1] >= rng[k][0]))
                return fail(SWARM_E_INVALID, "domain randomisation range %d must satisfy 0 < min <= max", k);
        if (c->dr_thrust_noise_std < 0 || c->dr_position_noise_std < 0 || c->dr_velocity_noise_std < 0 ||
            c->dr_obstacle_distance_noise_std < 0)
            return fail(SWARM_E_INVALID, "domain randomisation noise std must be >= 0");
        if (c->dr_delay_count < 0 || c->dr_delay_count > 4)
            return fail(SWARM_E_INVALID, "dr_delay_count must be in [0, 4]");
        for (int k = 0; k < c->dr_delay_count; ++k)
            if (c->dr_delay_values[k] < 0 || c->dr_delay_values[k] > 8 || !(c->dr_delay_probs[k] >= 0.0))
                return fail(SWARM_E_INVALID, "control delay %d: value must be in [0, 8] and its probability >= 0", k);
    }
    if ((int64_t)c->num_envs * c->num_drones > (int64_t)1 << 30)
        return fail(SWARM_E_UNSUPPORTED, "num_envs * num_drones too large");
    return SWARM_OK;
}

// largest float32 <= x: a float32-valued Python float d satisfies (d <= x) in double iff d <= this
float f32_floor_of(double x) {
    float f = (float)x;
    if ((double)f > x) f = std::nextafterf(f, -INFINITY);
    return f;
}

// Half-normal quantiles q[m] = Phi^-1(0.5 + (m + 0.5) / 512): the DR noise is a 9-bit table lookup (sign + m), so the
// CUDA path and the C oracle produce identical noise without depending on libm / MUFU rounding.
double inv_norm_cdf(double p) {  // Acklam's rational approximation
    static const double a[] = {-3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02,
                               1.383577518672690e+02, -3.066479806614716e+01, 2.506628277459239e+00};
    static const double b[] = {-5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02,
                               6.680131188771972e+01, -1.328068155288572e+01};
    static const double c[] = {-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00,
                               -2.549732539343734e+00, 4.374664141464968e+00, 2.938163982698783e+00};
    static const double d[] = {7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00,
                               3.754408661907416e+00};
    const double plow = 0.02425;
    if (p < plow) {
        const double q = std::sqrt(-2.0 * std::log(p));
        return (((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
               ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1.0);
    }
    if (p > 1.0 - plow) {
        const double q = std::sqrt(-2.0 * std::log(1.0 - p));
        return -(((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
               ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1.0);
    }
    const double q = p - 0.5, r = q * q;
    return (((((a[0] * r + a[1]) * r + a[2]) * r + a[3]) * r + a[4]) * r + a[5]) * q /
           (((((b[0] * r + b[1]) * r + b[2]) * r + b[3]) * r + b[4]) * r + 1.0);
}

void build_qtable(float* out) {
    for (int k = 0; k < 256; ++k) out[k] = (float)inv_norm_cdf(0.5 + ((double)k + 0.5) / 512.0);
}

void build_jump_table(int n_draws, std::vector<JumpEntry>& out) {
    typedef unsigned __int128 u128;
    const u128 mult = ((u128)2549297995355413924ULL << 64) | (u128)4865540595714422341ULL;
    out.resize((size_t)n_draws + 1);
    u128 a = 1, g = 0;  // A^0, G_0
    for (int k = 0; k <= n_draws; ++k) {
        out[(size_t)k].a_hi = (unsigned long long)(a >> 64);
        out[(size_t)k].a_lo = (unsigned long long)a;
        out[(size_t)k].g_hi = (unsigned long long)(g >> 64);
        out[(size_t)k].g_lo = (unsigned long long)g;
        g = g * mult + 1;  // G_{k+1} = A G_k + 1
        a = a * mult;
    }
}

void fill_params(const SwarmConfig& c, DevParams& p) {
    memset(&p, 0, sizeof(p));
    p.E = c.num_envs; p.N = c.num_drones; p.M = c.num_obstacles;
    p.K = c.env_kind != SWARM_KIND_SINGLE ? c.neighbor_k : 0;
    p.S = c.sensed_obstacles;
    p.D = obs_dim_of(c);
    p.R = 6 * p.N + 3;
    p.G = p.N <= 32 ? 32 / p.N : 1;
    p.nslots = (p.N + 31) / 32;
    p.n_tab = p.N <= 32 ? 32 : p.N;
    p.m_pad = p.M > 0 ? p.M : 1;
    p.n_draws = 3 * p.N + 3 + 3 * p.M;
    p.smem_per_warp = (2 * p.n_tab + p.G + p.G * p.m_pad) * 16 + 32 * p.D * 4;
    if (p.N <= 32) {
        // distance-matrix rows: stride >= N - 1, = 1 (mod 4) so row reads are conflict-free and the
        // symmetric writes are at most 2-way conflicted; the matrix aliases the obs staging tile
        int srow = ((p.N - 1 + 7) & ~7) > 1 ? ((p.N - 1 + 7) & ~7) : 1;  // rows are padded to whole blocks of 8
        while ((srow & 3) != 1) ++srow;
        p.srow = srow;
        const int region = 32 * p.D > 32 * srow + 8 ? 32 * p.D : 32 * srow + 8;
        // inbox: pos4[32] vel4[32] goal4[G] obst4[G*m_pad] actions[96 f32] step_count[G] ep_return[G]
        // (+ 32 B of per-env DR constants per env when domain randomisation is on)
        p.inbox_bytes = ((64 + p.G + p.G * p.m_pad) * 16 + 384 + 8 * p.G + 15) & ~15;
        if (c.dr_enabled) p.inbox_bytes += 32 * p.G;
        p.smem_per_warp = 2 * p.inbox_bytes + ((region + 3) & ~3) * 4;
    }
    p.dr_enabled = c.dr_enabled ? 1 : 0;
    if (p.dr_enabled) {
        p.dr_key0 = (unsigned)(c.dr_seed & 0xffffffffu);
        p.dr_key1 = (unsigned)(c.dr_seed >> 32);
        for (int r = 0; r < 10; ++r) {
            p.dr_rk0[r] = p.dr_key0 + (unsigned)r * 0x9E3779B9u;
            p.dr_rk1[r] = p.dr_key1 + (unsigned)r * 0xBB67AE85u;
        }
        p.env_index_base = c.env_index_base;
        const double* rng[6] = {c.dr_mass_scale, c.dr_max_accel_scale, c.dr_max_speed_scale, c.dr_dt_scale,
                                c.dr_obstacle_radius_scale, c.dr_world_size_scale};
        for (int k = 0; k < 6; ++k) { p.dr_lo[k] = rng[k][0]; p.dr_span[k] = rng[k][1] - rng[k][0]; }
        p.dr_max_accel = c.max_accel; p.dr_max_speed = c.max_speed; p.dr_dt = c.dt; p.dr_world = c.world_size;
        p.dr_r_c = c.collision_radius; p.dr_r_o = c.obstacle_radius;
        p.dr_std_thrust = (float)c.dr_thrust_noise_std; p.dr_std_pos = (float)c.dr_position_noise_std;
        p.dr_std_vel = (float)c.dr_velocity_noise_std; p.dr_std_obst = (float)c.dr_obstacle_distance_noise_std;
        p.dr_delay_hist = delay_hist_of(c);
        p.dr_delay_count = p.dr_delay_hist > 0 ? c.dr_delay_count : 0;
        double cum = 0.0;
        for (int k = 0; k < p.dr_delay_count; ++k) {
            cum += c.dr_delay_probs[k];
            p.dr_delay_values[k] = c.dr_delay_values[k];
            p.dr_delay_cum[k] = cum;
        }
    }
    p.n_others = (double)(p.N - 1);
    p.inv_n_others = p.N > 1 ? 1.0 / (double)(p.N - 1) : 0.0;
    p.max_steps = c.max_steps;
    // float32 constants (numpy NEP 50: a Python float meeting a float32 array / scalar is cast to f32)
    p.amax = (float)c.max_accel;                       // drone_swarm_env.py:107
    p.dt = (float)c.dt;                                // :109, :111
    p.vmax = (float)c.max_speed;                       // :181, :183
    p.eps_speed = (float)1e-8;                         // :181
    p.bound = (float)(c.world_size / 2.0);             // :113-117
    p.thr_goal = f32_floor_of(c.goal_radius);          // :124-127 Python-float compare
    p.thr_obst = (float)(c.collision_radius + c.obstacle_radius);  // :196-197
    p.thr_pair = (float)(2.0 * c.collision_radius);    // :205
    p.k_p = c.reward_progress_scale; p.r_goal = c.reward_goal; p.r_col = c.reward_collision;
    p.neg_k_f = -c.reward_formation_scale; p.d_star = c.desired_spacing;
    const double bound = c.world_size / 2.0;           // :71
    p.rng_lo = -bound; p.rng_range = bound - (-bound); // Generator.uniform(low, high): high - low
    if (c.env_kind == SWARM_KIND_PHYSICS) {
        // drone_physics_env.py: 5 draws per drone (position, mass noise, damping noise), 3 per obstacle, 4 for
        // the goal (:207-242); 24 sub-steps of 1/240 s (:323); g_comp 9.5 (:343) against gravity 9.81 (:197);
        // contacts as a point mass: ground at the URDF box's half height, spheres of radius 0.15 for the drones
        p.n_draws = 5 * p.N + 3 * p.M + 4;
        p.phys_substeps = (int)(c.dt * 240.0);
        p.phys_h = (float)(1.0 / 240.0);
        p.phys_g_net = (float)(9.5 - 9.81);
        p.phys_ground_z = 0.025f;
        p.thr_obst = (float)(c.obstacle_radius + 0.15);
        p.thr_pair = (float)(2.0 * 0.15);
        p.goal_radius_d = c.goal_radius;
        // domain randomisation on top: the sub-step length and the drone's contact radius take the place of dt / r_c
        p.dr_dt = 1.0 / 240.0;
        p.dr_r_c = 0.15;
    }
}

int bind_buffers(const SwarmHandle* h, const SwarmBuffers* b, DevParams& p) {
    if (!b) return fail(SWARM_E_NULL, "buffers is NULL");
    p = h->base;
    if (!b->pos4 || !b->vel4 || !b->goal4 || !b->obst4 || !b->step_count || !b->rng || !b->ep_return || !b->obs ||
        !b->reward || !b->dist || !b->terminated || !b->truncated || !b->reached || !b->collision ||
        !b->obs_valid || !b->all_terminated || !b->all_truncated)
        return fail(SWARM_E_NULL, "a required SwarmBuffers pointer is NULL");
    if ((reinterpret_cast<uintptr_t>(b->pos4) | reinterpret_cast<uintptr_t>(b->vel4) |
         reinterpret_cast<uintptr_t>(b->goal4) | reinterpret_cast<uintptr_t>(b->obst4) |
         reinterpret_cast<uintptr_t>(b->obs)) & 15u)
        return fail(SWARM_E_INVALID, "pos4 / vel4 / goal4 / obst4 / obs must be 16-byte aligned");
    p.pos4 = reinterpret_cast<float4*>(b->pos4); p.vel4 = reinterpret_cast<float4*>(b->vel4);
    p.goal4 = reinterpret_cast<float4*>(b->goal4); p.obst4 = reinterpret_cast<float4*>(b->obst4);
    p.step_count = b->step_count; p.rng = reinterpret_cast<unsigned long long*>(b->rng);
    p.ep_return = b->ep_return;
    p.obs = b->obs; p.reward = b->reward; p.reward64 = b->reward64; p.dist = b->dist;
    p.terminated = b->terminated; p.truncated = b->truncated; p.reached = b->reached;
    p.collision = b->collision; p.obs_valid = b->obs_valid;
    p.all_term = b->all_terminated; p.all_trunc = b->all_truncated;
    p.gs = b->global_state; p.episode_return = b->episode_return; p.episode_length = b->episode_length;
    p.stats = reinterpret_cast<unsigned long long*>(b->stats);
    p.jump = h->jump_dev;
    p.reset_mask = h->reset_mask_dev;
    if (p.dr_enabled) {
        if (!b->dr_params) return fail(SWARM_E_NULL, "dr_params is required when dr_enabled");
        if (reinterpret_cast<uintptr_t>(b->dr_params) & 15u) return fail(SWARM_E_INVALID, "dr_params must be 16-byte aligned");
        p.dr_params = reinterpret_cast<float4*>(b->dr_params);
        p.dr_qtable = h->qtable_dev;
        if (p.dr_delay_hist > 0) {
            if (!b->act_hist) return fail(SWARM_E_NULL, "act_hist is required when a control delay is configured");
            p.act_hist = b->act_hist;
        }
    }
    return SWARM_OK;
}

// The rotation-pass step kernel covers the BASELINE swarm shapes; its float64 formation sum is
// order-free only while every term is a multiple of 2^-37 and the sum stays below 2^16.
bool rot_eligible(const SwarmConfig& c) {
    if (c.env_kind != SWARM_KIND_SWARM || c.norm_mode != 0) return false;
    if (c.num_drones != 8 && c.num_drones != 16 && c.num_drones != 32) return false;
    if (c.neighbor_k != 3 || c.sensed_obstacles != 4) return false;
    if (c.num_obstacles < 4 || c.num_obstacles > 32 || (c.num_obstacles & 3)) return false;
    const double ds = c.desired_spacing;
    if (!(ds >= 0.0) || !(ds < 1024.0) || std::ldexp(ds, 37) != std::floor(std::ldexp(ds, 37))) return false;
    const double wscale = c.dr_enabled ? c.dr_world_size_scale[1] : 1.0;
    if (!(c.world_size * wscale < 1024.0)) return false;   // every distance < 2048
    const char* off = std::getenv("SWARM_B200_NO_ROT");
    return !(off && off[0] == '1');
}

// The wide rotation-pass kernels (several drones per lane) cover the dense swarms of BASELINE config 5.
bool rotx_eligible(const SwarmConfig& c) {
    if (c.env_kind != SWARM_KIND_SWARM || c.norm_mode != 0) return false;
    if (c.num_drones != 64 && c.num_drones != 128) return false;
    if (c.neighbor_k != 3 || c.sensed_obstacles != 4) return false;
    if (c.num_obstacles < 4 || c.num_obstacles > 32 || (c.num_obstacles & 3)) return false;
    const double ds = c.desired_spacing;
    if (!(ds >= 0.0) || !(ds < 512.0) || std::ldexp(ds, 37) != std::floor(std::ldexp(ds, 37))) return false;
    const double wscale = c.dr_enabled ? c.dr_world_size_scale[1] : 1.0;
    if (!(c.world_size * wscale < 256.0)) return false;   // 127 terms, each < 512: the float64 formation sum stays exact
    const char* off = std::getenv("SWARM_B200_NO_ROT");
    return !(off && off[0] == '1');
}

int launch(SwarmHandle* h, DevParams& p, int env_begin, int env_count, cudaStream_t stream, int slot = 0) {
    p.env_begin = env_begin;
    p.env_count = env_count;
    p.n_groups = (env_count + p.G - 1) / p.G;
    const bool rot = p.mode == kModeStep && h->rot_ok && (reinterpret_cast<uintptr_t>(p.actions) & 15u) == 0 &&
                     env_begin % p.G == 0;
    const int ctas_needed = (p.n_groups + kWarpsPerCta - 1) / kWarpsPerCta;
    const int resident = h->num_sms * h->blocks_per_sm;
    const int grid = ctas_needed < resident ? ctas_needed : resident;
    const int rot_needed = (p.n_groups + rot_warps_per_cta(p) - 1) / rot_warps_per_cta(p);
    const int rot_resident = h->num_sms * h->rot_blocks_per_sm;
    const int rot_grid = rot_needed < rot_resident ? rot_needed : rot_resident;
    // the wide kernel runs one env per warp with ~14 k instructions per item; SWARM_B200_ROTX_MIN_ENVS keeps launches
    // below that many envs on the general kernel (default 0: since the masked pass removed its serial exact-path
    // items the wide kernel is 2-3 x faster at small batches too)
    const bool rotx = p.mode == kModeStep && h->rotx_ok && (reinterpret_cast<uintptr_t>(p.actions) & 15u) == 0 &&
                      env_count >= h->rotx_min_envs;
    const int rotx_needed = (p.n_groups + rotx_warps_per_cta() - 1) / rotx_warps_per_cta();
    const int rotx_resident = h->num_sms * h->rotx_blocks_per_sm;
    const int rotx_grid = rotx_needed < rotx_resident ? rotx_needed : rotx_resident;
    if (p.n_steps < 1) p.n_steps = 1;
    p.fused_reset = rot && ((p.auto_reset && h->rot_fused) || p.n_steps > 1) ? 1 : 0;
    const bool two_launch = p.mode == kModeStep && p.auto_reset && (p.N <= 32 || rotx) && !p.fused_reset;
    // the step kernel lists the groups that need a reset for the launch behind it: two lists used alternately, the
    // current one chosen on the device (reset_epoch), so there is no per-launch host state (CUDA-graph safe)
    p.reset_count = h->reset_count_dev + 2 * slot;
    p.reset_epoch = h->reset_epoch_dev + slot;
    p.reset_list = h->reset_list_dev + env_begin / p.G;
    p.reset_list_stride = h->reset_list_stride;
    p.work_counter = h->work_counter_dev + 4 * slot;  // {step queue, warps done, reset queue, warps done}
    // NVTX ranges around the launches (a few ns without a profiler attached): nsys / ncu timelines show which API call
    // a kernel belongs to
    static const char* const kRange[] = {"swarm_step", "swarm_reset", "swarm_observe", "swarm_auto_reset"};
    nvtxRangePushA(kRange[p.mode & 3]);
    cudaError_t lerr;
    if (rot) lerr = launch_rot_kernel(p, rot_grid, stream);
    else if (rotx) lerr = launch_rotx_kernel(p, rotx_grid, stream);
    else lerr = launch_env_kernel(p, h->cfg.norm_mode, h->cfg.env_kind, grid, h->smem_bytes, stream);
    nvtxRangePop();
    CUDA_TRY(lerr);
    h->launches++;
    if (two_launch) {
        // N <= 32: the auto-reset runs as a second, tiny launch (kept out of the step kernel so each
        // launch's instruction working set fits the SM instruction cache)
        DevParams q = p;
        q.mode = kModeAutoReset;
        q.env_mask = h->reset_mask_dev;
        nvtxRangePushA(kRange[3]);
        if (rot) lerr = launch_rot_kernel(q, rot_grid, stream);
        else if (rotx) {
            const int rr = h->num_sms * h->rotx_reset_blocks_per_sm;
            lerr = launch_rotx_kernel(q, rotx_needed < rr ? rotx_needed : rr, stream);
        }
        else lerr = launch_env_kernel(q, h->cfg.norm_mode, h->cfg.env_kind, grid, h->smem_bytes, stream);
        nvtxRangePop();
        CUDA_TRY(lerr);
        h->launches++;
    }
    return SWARM_OK;
}

struct DeviceGuard {
    int prev = -1;
    bool active = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) {
            if (cudaSetDevice(dev) == cudaSuccess) active = true;
        }
    }
    ~DeviceGuard() { if (active) cudaSetDevice(prev); }
};

}  // namespace

extern "C" {

int swarm_abi_version(void) { return SWARM_ABI_VERSION; }

const char* swarm_last_error(void) { return g_err; }

int swarm_query_sizes(const SwarmConfig* cfg, SwarmSizes* out) {
    int rc = validate(cfg);
    if (rc != SWARM_OK) return rc;
    if (!out) return fail(SWARM_E_NULL, "out is NULL");
    const int64_t E = cfg->num_envs, N = cfg->num_drones, M = cfg->num_obstacles;
    const int64_t D = obs_dim_of(*cfg);
    out->obs_dim = D;
    out->state_dim = 6 * N + 3;
    out->pos4 = E * N * 4; out->vel4 = E * N * 4; out->goal4 = E * 4;
    out->obst4 = E * M * 4 > 4 ? E * M * 4 : 4;
    out->step_count = E; out->rng = E * 4; out->ep_return = E;
    out->actions = E * N * 3; out->obs = E * N * D; out->per_agent = E * N; out->per_env = E;
    out->global_state = E * (6 * N + 3);
    out->stats = SWARM_STATS_WORDS;
    out->dr_params = E * 8;
    out->act_hist = E * delay_hist_of(*cfg) * N * 3;
    return SWARM_OK;
}

int swarm_create(const SwarmConfig* cfg, SwarmHandle** out) {
    int rc = validate(cfg);
    if (rc != SWARM_OK) return rc;
    if (!out) return fail(SWARM_E_NULL, "out is NULL");
    *out = nullptr;
    int dev = cfg->device;
    if (dev < 0) CUDA_TRY(cudaGetDevice(&dev));
    DeviceGuard guard(dev);
    SwarmHandle* h = new (std::nothrow) SwarmHandle();
    if (!h) return fail(SWARM_E_INVALID, "out of host memory");
    h->cfg = *cfg;
    h->device = dev;
    h->launches = 0;
    h->jump_dev = nullptr;
    h->reset_mask_dev = nullptr;
    h->reset_count_dev = nullptr;
    h->reset_epoch_dev = nullptr;
    h->reset_list_dev = nullptr;
    h->work_counter_dev = nullptr;
    h->qtable_dev = nullptr;
    h->actions_dev = nullptr;
    h->flags_dev = nullptr;
    h->host_path_ready = false;
    fill_params(*cfg, h->base);
    h->smem_bytes = (size_t)h->base.smem_per_warp * kWarpsPerCta + (size_t)kWarpsPerCta * SWARM_STATS_WORDS * 8;

    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) { delete h; return fail(SWARM_E_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e)); }
    h->num_sms = prop.multiProcessorCount;
    if (h->smem_bytes > (size_t)prop.sharedMemPerBlockOptin) {
        const size_t need = h->smem_bytes;
        delete h;
        return fail(SWARM_E_UNSUPPORTED, "config needs %zu B shared memory per CTA (> %zu)", need,
                    (size_t)prop.sharedMemPerBlockOptin);
    }
    e = env_kernel_occupancy(h->base, cfg->norm_mode, cfg->env_kind, h->smem_bytes, &h->blocks_per_sm);
    if (e != cudaSuccess || h->blocks_per_sm < 1) {
        delete h;
        return fail(SWARM_E_CUDA, "kernel occupancy query failed: %s", cudaGetErrorString(e));
    }
    h->rot_ok = false;
    h->rot_blocks_per_sm = 0;
    if (rot_eligible(*cfg) && rot_smem_bytes(h->base) <= (size_t)prop.sharedMemPerBlockOptin) {
        e = rot_kernel_occupancy(h->base, &h->rot_blocks_per_sm);
        if (e != cudaSuccess) { delete h; return fail(SWARM_E_CUDA, "kernel occupancy query failed: %s", cudaGetErrorString(e)); }
        h->rot_ok = h->rot_blocks_per_sm >= 1;
    }
    h->rot_fused = false;   // (measured slower than the second launch so far: see DESIGN.md)
    h->multi_step_max_groups = 4 * h->num_sms * 28;   // up to ~4 groups per resident warp (measured: 16 384 groups of
                                                      // N = 32 one launch 46.7 us per step vs 49.6; 32 768: 84.4 vs 81.0)
    if (const char* mg = std::getenv("SWARM_B200_MULTI_STEP_MAX_GROUPS")) h->multi_step_max_groups = std::atoi(mg);
    if (const char* fr = std::getenv("SWARM_B200_FUSED_RESET")) h->rot_fused = fr[0] != '0';
    h->rotx_ok = false;
    h->rotx_blocks_per_sm = 0;
    h->rotx_reset_blocks_per_sm = 0;
    h->rotx_min_envs = 0;   // the wide kernel wins at every batch size now (128 envs: 46 vs 111 us per step at N = 128)
    if (const char* me = std::getenv("SWARM_B200_ROTX_MIN_ENVS")) h->rotx_min_envs = std::atoi(me);
    if (rotx_eligible(*cfg) && rotx_smem_bytes(h->base) <= (size_t)prop.sharedMemPerBlockOptin) {
        e = rotx_kernel_occupancy(h->base, &h->rotx_blocks_per_sm);
        if (e == cudaSuccess) {
            DevParams q = h->base;
            q.mode = kModeAutoReset;
            e = rotx_kernel_occupancy(q, &h->rotx_reset_blocks_per_sm);
        }
        if (e != cudaSuccess) { delete h; return fail(SWARM_E_CUDA, "kernel occupancy query failed: %s", cudaGetErrorString(e)); }
        h->rotx_ok = h->rotx_blocks_per_sm >= 1 && h->rotx_reset_blocks_per_sm >= 1;
    }
    std::vector<JumpEntry> table;
    build_jump_table(h->base.n_draws, table);
    e = cudaMalloc(&h->jump_dev, table.size() * sizeof(JumpEntry));
    if (e == cudaSuccess)
        e = cudaMemcpy(h->jump_dev, table.data(), table.size() * sizeof(JumpEntry), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&h->reset_mask_dev, (size_t)cfg->num_envs);
    if (e == cudaSuccess) e = cudaMemset(h->reset_mask_dev, 0, (size_t)cfg->num_envs);
    const size_t n_groups_all = ((size_t)cfg->num_envs + h->base.G - 1) / h->base.G;
    if (e == cudaSuccess) e = cudaMalloc(&h->reset_count_dev, sizeof(unsigned) * 2 * (kHostChunks + 1));
    if (e == cudaSuccess) e = cudaMemset(h->reset_count_dev, 0, sizeof(unsigned) * 2 * (kHostChunks + 1));
    h->reset_list_stride = (int)(n_groups_all + 1);
    if (e == cudaSuccess) e = cudaMalloc(&h->reset_list_dev, sizeof(int) * 2 * (n_groups_all + 1));
    if (e == cudaSuccess) e = cudaMalloc(&h->reset_epoch_dev, sizeof(unsigned) * (kHostChunks + 1));
    if (e == cudaSuccess) e = cudaMemset(h->reset_epoch_dev, 0, sizeof(unsigned) * (kHostChunks + 1));
    if (e == cudaSuccess) e = cudaMalloc(&h->work_counter_dev, sizeof(unsigned) * 4 * (kHostChunks + 1));
    if (e == cudaSuccess) e = cudaMemset(h->work_counter_dev, 0, sizeof(unsigned) * 4 * (kHostChunks + 1));
    if (e == cudaSuccess && cfg->dr_enabled) {
        std::vector<float> qt(256), qs(512);
        build_qtable(qt.data());
        for (int f = 0; f < 512; ++f) qs[f] = (f & 0x100) ? -qt[f & 0xFF] : qt[f & 0xFF];   // sign = bit 8 of the field
        e = cudaMalloc(&h->qtable_dev, 512 * sizeof(float));
        if (e == cudaSuccess) e = cudaMemcpy(h->qtable_dev, qs.data(), 512 * sizeof(float), cudaMemcpyHostToDevice);
    }
    if (e != cudaSuccess) {
        if (h->jump_dev) cudaFree(h->jump_dev);
        if (h->reset_mask_dev) cudaFree(h->reset_mask_dev);
        if (h->reset_count_dev) cudaFree(h->reset_count_dev);
        if (h->reset_epoch_dev) cudaFree(h->reset_epoch_dev);
        if (h->reset_list_dev) cudaFree(h->reset_list_dev);
        if (h->work_counter_dev) cudaFree(h->work_counter_dev);
        if (h->qtable_dev) cudaFree(h->qtable_dev);
        delete h;
        return fail(SWARM_E_CUDA, "jump table / reset mask allocation failed: %s", cudaGetErrorString(e));
    }
    *out = h;
    return SWARM_OK;
}

int swarm_destroy(SwarmHandle* h) {
    if (!h) return SWARM_OK;
    DeviceGuard guard(h->device);
    if (h->host_path_ready) {
        for (int c = 0; c < kHostChunks; ++c) {
            cudaStreamDestroy(h->chunk_stream[c]);
            cudaEventDestroy(h->chunk_done[c]);
        }
        cudaEventDestroy(h->caller_ready);
    }
    if (h->actions_dev) cudaFree(h->actions_dev);
    if (h->flags_dev) cudaFree(h->flags_dev);
    if (h->jump_dev) cudaFree(h->jump_dev);
    if (h->reset_mask_dev) cudaFree(h->reset_mask_dev);
    if (h->reset_count_dev) cudaFree(h->reset_count_dev);
    if (h->reset_epoch_dev) cudaFree(h->reset_epoch_dev);
    if (h->reset_list_dev) cudaFree(h->reset_list_dev);
    if (h->work_counter_dev) cudaFree(h->work_counter_dev);
    if (h->qtable_dev) cudaFree(h->qtable_dev);
    delete h;
    return SWARM_OK;
}

int swarm_seed(SwarmHandle* h, const SwarmBuffers* bufs, const uint64_t* seeds, const uint8_t* env_mask, void* stream) {
    if (!h) return fail(SWARM_E_NULL, "handle is NULL");
    if (!seeds) return fail(SWARM_E_NULL, "seeds is NULL");
    DeviceGuard guard(h->device);
    DevParams p;
    int rc = bind_buffers(h, bufs, p);
    if (rc != SWARM_OK) return rc;
    p.seeds = reinterpret_cast<const unsigned long long*>(seeds);
    p.env_mask = env_mask;
    CUDA_TRY(launch_seed_kernel(p, static_cast<cudaStream_t>(stream)));
    h->launches++;
    return SWARM_OK;
}

int swarm_reset(SwarmHandle* h, const SwarmBuffers* bufs, const uint8_t* env_mask, void* stream) {
    if (!h) return fail(SWARM_E_NULL, "handle is NULL");
    DeviceGuard guard(h->device);
    DevParams p;
    int rc = bind_buffers(h, bufs, p);
    if (rc != SWARM_OK) return rc;
    p.mode = kModeReset;
    p.env_mask = env_mask;
    return launch(h, p, 0, p.E, static_cast<cudaStream_t>(stream));
}

int swarm_observe(SwarmHandle* h, const SwarmBuffers* bufs, void* stream) {
    if (!h) return fail(SWARM_E_NULL, "handle is NULL");
    DeviceGuard guard(h->device);
    DevParams p;
    int rc = bind_buffers(h, bufs, p);
    if (rc != SWARM_OK) return rc;
    p.mode = kModeObserve;
    return launch(h, p, 0, p.E, static_cast<cudaStream_t>(stream));
}

int swarm_step(SwarmHandle* h, const SwarmBuffers* bufs, const float* actions, int auto_reset, void* stream) {
    if (!h) return fail(SWARM_E_NULL, "handle is NULL");
    if (!actions) return fail(SWARM_E_NULL, "actions is NULL");
    DeviceGuard guard(h->device);
    DevParams p;
    int rc = bind_buffers(h, bufs, p);
    if (rc != SWARM_OK) return rc;
    p.mode = kModeStep;
    p.auto_reset = auto_reset ? 1 : 0;
    p.actions = actions;
    return launch(h, p, 0, p.E, static_cast<cudaStream_t>(stream));
}

int swarm_step_many(SwarmHandle* h, const SwarmBuffers* bufs, const float* actions, int n_steps, int auto_reset,
                    void* stream) {
    if (!h) return fail(SWARM_E_NULL, "handle is NULL");
    if (!actions) return fail(SWARM_E_NULL, "actions is NULL");
    if (n_steps < 0) return fail(SWARM_E_INVALID, "n_steps must be >= 0");
    DeviceGuard guard(h->device);
    DevParams p;
    int rc = bind_buffers(h, bufs, p);
    if (rc != SWARM_OK) return rc;
    p.mode = kModeStep;
    p.auto_reset = auto_reset ? 1 : 0;
    const size_t per_step = (size_t)p.E * p.N * 3;
    // small batches are bound by launch latency and by the ramp / tail of every launch, not by the device: there ALL
    // the steps run in one launch of the fused rotation-pass kernel (static env ownership per warp, in-warp
    // auto-reset, no barrier between steps); larger ones are faster as one step + one reset launch per step
    const int n_groups_all = (p.E + p.G - 1) / p.G;
    if (n_steps > 1 && h->rot_ok && n_groups_all <= h->multi_step_max_groups &&
        (reinterpret_cast<uintptr_t>(actions) & 15u) == 0 && (per_step & 3u) == 0) {
        DevParams q = p;
        q.actions = actions;
        q.n_steps = n_steps;
        q.action_step_stride = (long long)per_step;
        return launch(h, q, 0, p.E, static_cast<cudaStream_t>(stream));
    }
    for (int t = 0; t < n_steps; ++t) {
        DevParams q = p;
        q.actions = actions + (size_t)t * per_step;
        rc = launch(h, q, 0, p.E, static_cast<cudaStream_t>(stream));
        if (rc != SWARM_OK) return rc;
    }
    return SWARM_OK;
}

int swarm_step_host(SwarmHandle* h, const SwarmBuffers* bufs, const float* actions_host, const SwarmHostOut* out,
                    int auto_reset, void* stream) {
    if (!h) return fail(SWARM_E_NULL, "handle is NULL");
    if (!actions_host || !out) return fail(SWARM_E_NULL, "actions_host / out_host is NULL");
    DeviceGuard guard(h->device);
    DevParams p;
    int rc = bind_buffers(h, bufs, p);
    if (rc != SWARM_OK) return rc;
    if (out->global_state && !p.gs) return fail(SWARM_E_INVALID, "host global_state requested but bufs->global_state is NULL");
    if (!h->host_path_ready) {
        for (int c = 0; c < kHostChunks; ++c) {
            CUDA_TRY(cudaStreamCreateWithFlags(&h->chunk_stream[c], cudaStreamNonBlocking));
            CUDA_TRY(cudaEventCreateWithFlags(&h->chunk_done[c], cudaEventDisableTiming));
        }
        CUDA_TRY(cudaEventCreateWithFlags(&h->caller_ready, cudaEventDisableTiming));
        CUDA_TRY(cudaMalloc(&h->actions_dev, (size_t)p.E * p.N * 3 * sizeof(float)));
        h->host_path_ready = true;
    }
    if (out->flags && out->block_bytes == 0 && !h->flags_dev)
        CUDA_TRY(cudaMalloc(&h->flags_dev, ((size_t)p.E * p.N + 3) & ~(size_t)3));
    // the internal chunk streams start behind whatever the caller has enqueued on `stream` (an event, not a
    // device-wide drain); the call returns after a host wait on every chunk, so later work on any stream is behind it
    CUDA_TRY(cudaEventRecord(h->caller_ready, static_cast<cudaStream_t>(stream)));
    p.mode = kModeStep;
    p.auto_reset = auto_reset ? 1 : 0;
    p.actions = h->actions_dev;
    const int E = p.E, N = p.N, D = p.D, R = p.R;
    // chunk boundaries are multiples of G so a warp's env group never straddles two launches
    int chunks = kHostChunks;
    int groups_total = (E + p.G - 1) / p.G;
    if (groups_total < chunks * 64 || out->block_bytes > 0) chunks = 1;
    const int groups_per_chunk = (groups_total + chunks - 1) / chunks;
    for (int c = 0; c < chunks; ++c) {
        const int e0 = c * groups_per_chunk * p.G;
        if (e0 >= E) break;
        const int e1 = (e0 + groups_per_chunk * p.G) < E ? (e0 + groups_per_chunk * p.G) : E;
        const size_t ne = (size_t)(e1 - e0);
        cudaStream_t s = h->chunk_stream[c];
        CUDA_TRY(cudaStreamWaitEvent(s, h->caller_ready, 0));
        CUDA_TRY(cudaMemcpyAsync(h->actions_dev + (size_t)e0 * N * 3, actions_host + (size_t)e0 * N * 3,
                                 ne * N * 3 * sizeof(float), cudaMemcpyHostToDevice, s));
        DevParams pc = p;
        rc = launch(h, pc, e0, e1 - e0, s, c + 1);
        if (rc != SWARM_OK) return rc;
 if (out->block_bytes > 0) {
            // small batches (the E = 1 facade envs): every output lives in ONE device block mirrored by one pinned
            // host block -- one copy instead of twelve
            CUDA_TRY(cudaMemcpyAsync(out->block_host, out->block_dev, (size_t)out->block_bytes, cudaMemcpyDeviceToHost, s));
            CUDA_TRY(cudaEventRecord(h->chunk_done[c], s));
            continue;
        }
#define D2H(field, devptr, per_env_elems, type)                                                         \
        if (out->field)                                                                                 \
            CUDA_TRY(cudaMemcpyAsync(out->field + (size_t)e0 * (per_env_elems), (devptr) + (size_t)e0 * (per_env_elems), \
                                     ne * (per_env_elems) * sizeof(type), cudaMemcpyDeviceToHost, s))
        D2H(obs, p.obs, (size_t)N * D, float);
        D2H(reward, p.reward, (size_t)N, float);
        if (out->reward64 && p.reward64) {
            CUDA_TRY(cudaMemcpyAsync(out->reward64 + (size_t)e0 * N, p.reward64 + (size_t)e0 * N, ne * N * sizeof(double),
                                     cudaMemcpyDeviceToHost, s));
        }
        D2H(dist, p.dist, (size_t)N, float);
        D2H(terminated, p.terminated, (size_t)N, uint8_t);
        D2H(truncated, p.truncated, (size_t)N, uint8_t);
        D2H(reached, p.reached, (size_t)N, uint8_t);
        D2H(collision, p.collision, (size_t)N, uint8_t);
        D2H(obs_valid, p.obs_valid, (size_t)N, uint8_t);
        if (out->flags) {
            const long long a_first = (long long)e0 * N, n_ag = (long long)ne * N;
            const bool vec = ((a_first | n_ag) & 3) == 0;
            const long long threads = vec ? n_ag / 4 : n_ag;
            swarm_pack_flags_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(
                p.terminated, p.truncated, p.reached, p.collision, p.obs_valid, h->flags_dev, a_first, n_ag);
            CUDA_TRY(cudaGetLastError());
            ++h->launches;
            CUDA_TRY(cudaMemcpyAsync(out->flags + a_first, h->flags_dev + a_first, (size_t)n_ag, cudaMemcpyDeviceToHost, s));
        }
        D2H(all_terminated, p.all_term, (size_t)1, uint8_t);
        D2H(all_truncated, p.all_trunc, (size_t)1, uint8_t);
        D2H(global_state, p.gs, (size_t)R, float);
#undef D2H
        CUDA_TRY(cudaEventRecord(h->chunk_done[c], s));
    }
    for (int c = 0; c < chunks; ++c) {
        if (c * groups_per_chunk * p.G >= E) break;
        CUDA_TRY(cudaEventSynchronize(h->chunk_done[c]));
    }
    return SWARM_OK;
}

int64_t swarm_launch_count(const SwarmHandle* h) { return h ? h->launches : 0; }

int swarm_dr_quantile_table(float* out256) {
    if (!out256) return fail(SWARM_E_NULL, "out is NULL");
    build_qtable(out256);
    return SWARM_OK;
}

}  // extern "C"
