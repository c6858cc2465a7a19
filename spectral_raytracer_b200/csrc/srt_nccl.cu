// srt_nccl.cu -- srt_reduce(): sum the spectral accumulation buffers of n contexts (one per device, one
// host process) into ctxs[0] with NCCL over NVLink.  Built as libsrt_nccl.so on top of libsrt.so's public
// C ABI only, so a host without NCCL can still use libsrt.so.  (The one-process-per-GPU path of bench.py uses
// torch.distributed on the same buffers instead, see spectral_raytracer_b200/distributed.py.)
//
// The reduce is the only exchange step of the render path: frames are sharded, every context
// holds the radiance sum of its own frames, the image is the sum divided by the total frame count.
//
// Communicators are created once per set of devices and kept (ncclCommInitAll costs hundreds of milliseconds,
// a reduce of the 265 MB buffer well under one); srt_reduce_shutdown() destroys them.
#include <cuda_runtime.h>
#include <nccl.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/srt.h"

namespace {
thread_local std::string g_err;
thread_local float g_last_ms = 0.0f;

std::mutex g_mu;
std::map<std::vector<int>, std::vector<ncclComm_t>> g_comms;  // device list (in rank order) -> communicators

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

// the communicators of this device list, created on first use
int comms_for(const std::vector<int>& devs, std::vector<ncclComm_t>** out) {
    auto it = g_comms.find(devs);
    if (it == g_comms.end()) {
        std::vector<ncclComm_t> comms(devs.size());
        ncclResult_t r = ncclCommInitAll(comms.data(), (int)devs.size(), devs.data());
        if (r != ncclSuccess) return fail(SRT_ERR_CUDA, std::string("ncclCommInitAll: ") + ncclGetErrorString(r));
        it = g_comms.emplace(devs, std::move(comms)).first;
    }
    *out = &it->second;
    return SRT_OK;
}
}  // namespace

extern "C" {

const char* srt_reduce_last_error(void) { return g_err.c_str(); }

float srt_reduce_last_ms(void) { return g_last_ms; }

void srt_reduce_shutdown(void) {
    std::lock_guard<std::mutex> lock(g_mu);
    for (auto& kv : g_comms)
        for (ncclComm_t c : kv.second) ncclCommDestroy(c);
    g_comms.clear();
}

int srt_reduce(srt_ctx* const* ctxs, uint32_t n) {
    try {
        if (!ctxs || n == 0) return fail(SRT_ERR_INVALID_ARGUMENT, "srt_reduce: no contexts");
        std::vector<int> devs(n);
        std::vector<float*> bufs(n);
        std::vector<cudaStream_t> streams(n);
        size_t count = 0;
        uint64_t frames = 0;
        for (uint32_t i = 0; i < n; ++i) {
            if (!ctxs[i]) return fail(SRT_ERR_INVALID_ARGUMENT, "srt_reduce: null context");
            size_t c = 0;
            bufs[i] = static_cast<float*>(srt_accum_device_ptr(ctxs[i], &c));
            if (i == 0) count = c;
            if (c != count) return fail(SRT_ERR_INVALID_ARGUMENT, "srt_reduce: contexts have different image sizes / spectral widths");
            devs[i] = srt_device(ctxs[i]);
            for (uint32_t j = 0; j < i; ++j)
                if (devs[j] == devs[i])
                    return fail(SRT_ERR_INVALID_ARGUMENT, "srt_reduce: two contexts on the same device (NCCL needs one rank per device)");
            streams[i] = static_cast<cudaStream_t>(srt_stream(ctxs[i]));
            frames += srt_frames_accumulated(ctxs[i]);
        }
        g_last_ms = 0.0f;
        if (n == 1) return SRT_OK;
        std::lock_guard<std::mutex> lock(g_mu);
        std::vector<ncclComm_t>* comms = nullptr;
        int rc = comms_for(devs, &comms);
        if (rc) return rc;
        int prev = 0;
        cudaGetDevice(&prev);
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        cudaSetDevice(devs[0]);
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, streams[0]);
        ncclGroupStart();
        for (uint32_t i = 0; i < n; ++i) {
            cudaSetDevice(devs[i]);
            ncclResult_t r = ncclReduce(bufs[i], bufs[i], count, ncclFloat, ncclSum, 0, (*comms)[i], streams[i]);
            if (r != ncclSuccess) rc = fail(SRT_ERR_CUDA, std::string("ncclReduce: ") + ncclGetErrorString(r));
        }
        ncclResult_t r = ncclGroupEnd();
        if (r != ncclSuccess && rc == SRT_OK) rc = fail(SRT_ERR_CUDA, std::string("ncclGroupEnd: ") + ncclGetErrorString(r));
        cudaSetDevice(devs[0]);
        cudaEventRecord(e1, streams[0]);
        for (uint32_t i = 0; i < n; ++i) {
            cudaSetDevice(devs[i]);
            if (cudaStreamSynchronize(streams[i]) != cudaSuccess && rc == SRT_OK) rc = fail(SRT_ERR_CUDA, "srt_reduce: stream synchronize failed");
        }
        if (rc == SRT_OK) cudaEventElapsedTime(&g_last_ms, e0, e1);
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        cudaSetDevice(prev);
        if (rc != SRT_OK) return rc;
        // the sum now lives in ctxs[0]; the others start their next shard from an empty image, so that a second
        // round (progressive rendering, resume) does not add their old radiance again
        for (uint32_t i = 1; i < n; ++i) {
            rc = srt_clear(ctxs[i]);
            if (rc) return fail(rc, std::string("srt_reduce: ") + srt_last_error(ctxs[i]));
        }
        return srt_set_frames_accumulated(ctxs[0], frames);
    } catch (const std::exception& e) {
        return fail(SRT_ERR_CUDA, std::string("srt_reduce: ") + e.what());
    } catch (...) {
        return fail(SRT_ERR_CUDA, "srt_reduce: unknown exception");
    }
}

}  // extern "C"
