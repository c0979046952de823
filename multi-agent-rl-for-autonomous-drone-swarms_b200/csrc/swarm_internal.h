// swarm_internal.h -- shared between the kernel TU and the C-ABI TU (not installed).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "swarm_b200.h"

namespace swarm {

#ifndef SWARM_WARPS_PER_CTA
#define SWARM_WARPS_PER_CTA 8
#endif
#ifndef SWARM_MIN_BLOCKS
#define SWARM_MIN_BLOCKS 3  // register budget: 65536 / (3 * 256) = 85 per thread
#endif
constexpr int kWarpsPerCta = SWARM_WARPS_PER_CTA;
constexpr int kThreadsPerCta = kWarpsPerCta * 32;
constexpr int kMinBlocksPerSm = SWARM_MIN_BLOCKS;

enum Mode : int { kModeStep = 0, kModeReset = 1, kModeObserve = 2, kModeAutoReset = 3 };  // 3: reset that keeps the step's reward / flags

// One entry per draw count k: state_{n+k} = A^k * state_n + G_k * inc  (mod 2^128)
struct __align__(16) JumpEntry {
    unsigned long long a_hi, a_lo, g_hi, g_lo;
};

// Kernel parameters (passed by value; lives in the constant bank).
struct DevParams {
    // shapes
    int E, N, M, K, S, D, R;      // R = 6N + 3
    int G;                        // env instances per warp (N <= 32: 32 / N, else 1)
    int nslots;                   // drones per lane (ceil(N / 32))
    int n_tab;                    // float4 entries of the per-warp position / velocity tables
    int m_pad;                    // obstacle table entries per env (max(M, 1))
    int n_draws;                  // 3N + 3 + 3M uniform draws per reset
    int env_begin, env_count;     // env range of this launch
    int n_groups;                 // ceil(env_count / G)
    int smem_per_warp;            // bytes
    int srow;                     // N <= 32 kernel: row stride (floats) of the per-warp distance matrix
    int inbox_bytes;              // N <= 32 kernel: bytes of one input inbox (two per warp, then the staging region)
    int mode;                     // Mode
    int auto_reset;
    int fused_reset;              // rotation-pass step launch: the auto-reset runs inside it (one launch per step)
    int n_steps;                  // ... and that launch runs this many consecutive steps (swarm_step_many, small batches):
                                  //   step t reads actions + t * action_step_stride
    long long action_step_stride;
    int max_steps;
    // float32 constants the reference's numpy expressions effectively use
    float amax, dt, vmax, eps_speed, bound;
    float thr_goal;               // largest f32 <= goal_radius (double compare of an f32-valued float)
    float thr_obst, thr_pair;     // f32(r_c + r_o), f32(2 r_c)
    // Python-float (double) constants
    double k_p, r_goal, r_col, neg_k_f, d_star;
    double rng_lo, rng_range;     // uniform(-W/2, W/2): lo, hi - lo
    double n_others, inv_n_others;  // N - 1 and RN(1 / (N - 1)) for the mean of the formation errors
    // physics env (point-mass DronePhysicsEnv)
    double goal_radius_d;         // reached = (double)dist < goal_radius
    float phys_h, phys_g_net, phys_ground_z;  // 1/240 s, 9.5 - 9.81, drone half height
    int phys_substeps;            // int(dt * 240)
    // state
    float4* pos4; float4* vel4; float4* goal4; float4* obst4;
    int* step_count; unsigned long long* rng; float* ep_return;
    // inputs
    const float* actions; const uint8_t* env_mask; const unsigned long long* seeds;
    // outputs
    float* obs; float* reward; double* reward64; float* dist;
    uint8_t *terminated, *truncated, *reached, *collision, *obs_valid, *all_term, *all_trunc;
    float* gs; float* episode_return; int* episode_length;
    unsigned long long* stats;
    uint8_t* reset_mask;          // [E] N <= 32 step kernel -> aux kernel: env needs its auto-reset
    // auto-reset hand-over between the step launch and the reset launch behind it: two compacted lists of
    // groups with an env to reset, used alternately.  Which one is current is decided ON THE DEVICE by the parity of
    // `reset_epoch` (incremented by the last warp to leave an auto-reset step launch, which also zeroes the other
    // list's counter), so the host keeps no per-launch state and a captured CUDA graph can be replayed any number
    // of times.  Step launch t appends to list (epoch & 1); the reset launch behind it reads list ((epoch - 1) & 1).
    unsigned* reset_count;        // [2] entries on each list
    unsigned* reset_epoch;        // auto-reset step launches completed so far (this launch slot)
    int* reset_list;              // [2][reset_list_stride] first env of every group with an env to reset
    int reset_list_stride;
    unsigned* work_counter;       // rotation-pass step kernel: {next group, warps done} of this launch slot
    const JumpEntry* jump;        // [n_draws + 1]
    // ---- domain randomisation (N <= 32 kernels, norm_mode 0)
    int dr_enabled;
    unsigned dr_key0, dr_key1;    // Philox key = dr_seed
    unsigned dr_rk0[10], dr_rk1[10];  // its round keys: key + r * Weyl constant
    long long env_index_base;
    double dr_lo[6], dr_span[6];  // mass, max_accel, max_speed, dt, obstacle_radius, world_size: min, max - min
    double dr_max_accel, dr_max_speed, dr_dt, dr_world, dr_r_c, dr_r_o;
    float dr_std_thrust, dr_std_pos, dr_std_vel, dr_std_obst;
    float4* dr_params;            // [E][2] float4
    const float* dr_qtable;       // [512] signed normal quantiles: entry f = the normal of the 9-bit field f
    int dr_delay_count, dr_delay_hist;   // control delay: number of choices, ring size H (0 = off)
    int dr_delay_values[4];
    double dr_delay_cum[4];
    float* act_hist;              // [E][H][N][3] ring of submitted commands
};

// kernel selection (swarm_kernels.cu)
cudaError_t launch_env_kernel(const DevParams& p, int norm_mode, int env_kind, int grid, size_t smem_bytes,
                              cudaStream_t stream);
cudaError_t env_kernel_occupancy(const DevParams& p, int norm_mode, int env_kind, size_t smem_bytes,
                                 int* blocks_per_sm);
cudaError_t launch_seed_kernel(const DevParams& p, cudaStream_t stream);
// rotation-pass step kernel (swarm_step_rot.cu)
size_t rot_smem_bytes(const DevParams& p);
cudaError_t launch_rot_kernel(const DevParams& p, int grid, cudaStream_t stream);
cudaError_t rot_kernel_occupancy(const DevParams& p, int* blocks_per_sm);
int rot_warps_per_cta(const DevParams& p);
// rotation-pass kernels for N = 64 / 128 (swarm_step_rotx.cu)
size_t rotx_smem_bytes(const DevParams& p);
cudaError_t launch_rotx_kernel(const DevParams& p, int grid, cudaStream_t stream);
cudaError_t rotx_kernel_occupancy(const DevParams& p, int* blocks_per_sm);
int rotx_warps_per_cta();

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) costs ~1 us: do it once per (device, kernel), not per launch
// (small batches are launch-bound: 10-14 us of host time per step)
inline cudaError_t ensure_dynamic_smem(const void* kernel, size_t smem) {
    constexpr int kSlots = 128;
    static thread_local const void* done_kernel[kSlots];
    static thread_local size_t done_smem[kSlots];
    static thread_local int done_dev[kSlots];
    static thread_local int n_done = 0;
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    for (int k = 0; k < n_done; ++k)
        if (done_kernel[k] == kernel && done_dev[k] == dev && done_smem[k] >= smem) return cudaSuccess;
    err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err == cudaSuccess && n_done < kSlots) {
        done_kernel[n_done] = kernel; done_smem[n_done] = smem; done_dev[n_done] = dev;
        ++n_done;
    }
    return err;
}

}  // namespace swarm
