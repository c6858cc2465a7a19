// srt_nccl.cu -- srt_reduce(): sum the spectral accumulation buffers of n contexts (one per device, one
// host process) into ctxs[0] with NCCL over NVLink.  Built as libsrt_nccl.so on top of libsrt.so's public
// C ABI only, so a host without NCCL can still use libsrt.so.  (The one-process-per-GPU path of bench.py uses
// torch.distributed on the same buffers instead, see spectral_raytracer_b200/distributed.py.)
//
// The reduce is the only exchange step of the render path: frames are sharded, every context
// holds the radiance sum of its own frames, the image is the sum divided by the total frame count.
#include <cuda_runtime.h>
#include <nccl.h>

#include <string>
#include <vector>

#include "../../include/srt.h"

namespace {
thread_local std::string g_err;
}

extern "C" {

const char* srt_reduce_last_error(void) { return g_err.c_str(); }

int srt_reduce(srt_ctx* const* ctxs, uint32_t n) {
    if (!ctxs || n == 0) {
        g_err = "srt_reduce: no contexts";
        return SRT_ERR_INVALID_ARGUMENT;
    }
    std::vector<int> devs(n);
    std::vector<float*> bufs(n);
    std::vector<cudaStream_t> streams(n);
    size_t count = 0;
    uint64_t frames = 0;
    for (uint32_t i = 0; i < n; ++i) {
        if (!ctxs[i]) {
            g_err = "srt_reduce: null context";
            return SRT_ERR_INVALID_ARGUMENT;
        }
        size_t c = 0;
        bufs[i] = static_cast<float*>(srt_accum_device_ptr(ctxs[i], &c));
        if (i == 0) count = c;
        if (c != count) {
            g_err = "srt_reduce: contexts have different image sizes / spectral widths";
            return SRT_ERR_INVALID_ARGUMENT;
        }
        devs[i] = srt_device(ctxs[i]);
        for (uint32_t j = 0; j < i; ++j)
            if (devs[j] == devs[i]) {
                g_err = "srt_reduce: two contexts on the same device (NCCL needs one rank per device)";
                return SRT_ERR_INVALID_ARGUMENT;
            }
        streams[i] = static_cast<cudaStream_t>(srt_stream(ctxs[i]));
        frames += srt_frames_accumulated(ctxs[i]);
    }
    if (n == 1) return SRT_OK;
    int prev = 0;
    cudaGetDevice(&prev);
    std::vector<ncclComm_t> comms(n);
    ncclResult_t r = ncclCommInitAll(comms.data(), (int)n, devs.data());
    if (r != ncclSuccess) {
        g_err = std::string("ncclCommInitAll: ") + ncclGetErrorString(r);
        return SRT_ERR_CUDA;
    }
    int rc = SRT_OK;
    ncclGroupStart();
    for (uint32_t i = 0; i < n; ++i) {
        cudaSetDevice(devs[i]);
        r = ncclReduce(bufs[i], bufs[i], count, ncclFloat, ncclSum, 0, comms[i], streams[i]);
        if (r != ncclSuccess) {
            g_err = std::string("ncclReduce: ") + ncclGetErrorString(r);
            rc = SRT_ERR_CUDA;
        }
    }
    r = ncclGroupEnd();
    if (r != ncclSuccess && rc == SRT_OK) {
        g_err = std::string("ncclGroupEnd: ") + ncclGetErrorString(r);
        rc = SRT_ERR_CUDA;
    }
    for (uint32_t i = 0; i < n; ++i) {
        cudaSetDevice(devs[i]);
        if (cudaStreamSynchronize(streams[i]) != cudaSuccess && rc == SRT_OK) {
            g_err = "srt_reduce: stream synchronize failed";
            rc = SRT_ERR_CUDA;
        }
        ncclCommDestroy(comms[i]);
    }
    cudaSetDevice(prev);
    if (rc == SRT_OK) srt_set_frames_accumulated(ctxs[0], frames);
    return rc;
}

}  // extern "C"
