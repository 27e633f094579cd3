/* torchrl_b200.h -- C ABI of libtorchrl_b200.so (sm_100a kernels for the torchrl hot path).
 *
 * The reference (RchalYang/torchrl, pure Python) has no FFI of its own; these are the
 * entry points a ctypes binding of the reference would call in place of the Python/NumPy
 * code cited beside each declaration (paths relative to the reference root).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; the caller owns
 *     all memory (nothing is allocated or retained by the library);
 *   - arrays are dense, row-major, time-major: (T, N, D) with the env index N contiguous
 *     over D -- the layout of the reference's buffers (replay_buffers/base.py:19-29);
 *   - float data are fp32; 0/1 flags (terminals, time_limits) are uint8;
 *   - `stream` is a cudaStream_t passed as void*; launches are asynchronous, no host sync;
 *   - return value: 0 = ok, <0 = argument error (TRL_E*), >0 = cudaError_t of the launch;
 *     trl_last_error() returns the message of the calling thread's last failure.
 */
#ifndef TORCHRL_B200_H
#define TORCHRL_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TRL_OK 0
#define TRL_EINVAL (-1)
#define TRL_EALIGN (-2)
#define TRL_EUNSUPPORTED (-3)

const char* trl_last_error(void);
int trl_abi_version(void);
int trl_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* K6 -- replaces OnPolicyReplayBufferBase.generalized_advantage_estimation
 * (torchrl/replay_buffers/on_policy.py:16-44).  variant: 1 = auto (default),
 * 0 = serial-in-T reference-shaped kernel, 2 = force 4-env vectorised, 3 = force scalar. */
int trl_gae_scan(const float* rewards, const float* values, const uint8_t* terminals,
                 const uint8_t* time_limits, const float* last_value, float* advs, float* returns,
                 int64_t T, int64_t N, float gamma, float tau, int time_limit_filter, int variant,
                 void* stream);

/* K6 -- replaces OnPolicyReplayBufferBase.discount_reward (torchrl/replay_buffers/on_policy.py:46-70). */
int trl_discount_return(const float* rewards, const float* values, const uint8_t* terminals,
                        const uint8_t* time_limits, const float* last_value, float* advs, float* returns,
                        int64_t T, int64_t N, float gamma, int time_limit_filter, int variant,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TORCHRL_B200_H */
