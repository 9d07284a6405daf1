// rg_peer.inl -- gather of per-shard [cost | grad] blocks through peer memory (included by rg_api.cu).
//
// north_star: "an NCCL allgather over NVLink of per-shard costs and gradients". Measured: with the
// evaluation kernels filling every SM, NCCL's all-gather kernel has to win SM slots on all ranks at once and took longer
// than the 1024-pulse shard's kernels. Here every rank pushes its block straight into the peers' buffers (CUDA IPC
// mappings, NVLink/NVSwitch peer stores): by the copy engines (mode 0, no SM at all) or by one short store kernel (mode 1).
// No counterpart in the reference (single CPU process).

struct PeerPtrs { double* p[16]; };

static __global__ void __launch_bounds__(256) k_peer_scatter16(const double2* __restrict__ src, size_t n2, PeerPtrs pp,
                                                               int np, size_t off2) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
        const double2 v = __ldcs(src + i);
#pragma unroll 1
        for (int q = 0; q < np; q++) reinterpret_cast<double2*>(pp.p[q])[off2 + i] = v;
    }
}
static __global__ void __launch_bounds__(256) k_peer_scatter8(const double* __restrict__ src, size_t n, PeerPtrs pp, int np,
                                                              size_t off) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double v = src[i];
        for (int q = 0; q < np; q++) pp.p[q][off + i] = v;
    }
}

static int peer_streams(rg_ctx* c) {
    if (c->s_peer[0]) return RG_OK;
    CU(c, cudaSetDevice(c->device));
    for (int i = 0; i < RG_PEER_STREAMS; i++) {
        CU(c, cudaStreamCreateWithFlags(&c->s_peer[i], cudaStreamNonBlocking));
        CU(c, cudaEventCreateWithFlags(&c->ev_peer_join[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < 2; i++) CU(c, cudaEventCreateWithFlags(&c->ev_gather[i], cudaEventDisableTiming));
    CU(c, cudaEventCreateWithFlags(&c->ev_src, cudaEventDisableTiming));
    return RG_OK;
}

extern "C" int rg_peer_buffer_create(rg_ctx* c, uint64_t bytes, void** dptr, unsigned char handle[64]) {
    if (!c || !dptr || !handle || bytes == 0) return RG_ERR_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle is 64 bytes");
    CU(c, cudaSetDevice(c->device));
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); RG_FAIL(c, RG_ERR_NOMEM, "peer buffer allocation failed"); }
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); cudaGetLastError(); RG_FAIL(c, RG_ERR_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); }
    memcpy(handle, &h, 64);
    *dptr = p;
    return RG_OK;
}

extern "C" int rg_peer_buffer_open(rg_ctx* c, const unsigned char handle[64], void** dptr) {
    if (!c || !dptr || !handle) return RG_ERR_INVALID;
    CU(c, cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { cudaGetLastError(); RG_FAIL(c, RG_ERR_CUDA, "cudaIpcOpenMemHandle: %s", cudaGetErrorString(e)); }
    *dptr = p;
    return RG_OK;
}

extern "C" int rg_peer_buffer_close(rg_ctx* c, void* dptr) {
    if (!c) return RG_ERR_INVALID;
    if (!dptr) return RG_OK;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaIpcCloseMemHandle(dptr));
    return RG_OK;
}

extern "C" int rg_peer_buffer_destroy(rg_ctx* c, void* dptr) {
    if (!c) return RG_ERR_INVALID;
    if (!dptr) return RG_OK;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaFree(dptr));
    return RG_OK;
}

extern "C" int rg_gather_to_peers(rg_ctx* c, const void* src, uint64_t bytes, int32_t npeers, void* const* peer_base,
                                  uint64_t dst_offset, int32_t slot, int32_t mode) {
    if (!c) return RG_ERR_INVALID;
    if (npeers < 0 || npeers > 16 || slot < 0 || slot > 1 || (npeers > 0 && (!src || !peer_base)) || (bytes & 7) || (dst_offset & 7))
        RG_FAIL(c, RG_ERR_INVALID, "bad gather arguments (at most 16 peers, slot 0/1, sizes in whole doubles)");
    if (peer_streams(c) != RG_OK) return RG_ERR_CUDA;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaEventRecord(c->ev_src, c->stream));                      // the source block is complete here
    cudaStream_t s0 = c->s_peer[0];
    CU(c, cudaStreamWaitEvent(s0, c->ev_src, 0));
    if (bytes && npeers) {
        if (mode == 0) {
            // One copy per peer, the peers spread over RG_PEER_STREAMS streams (copy engines).  Cutting a block into pieces was
            // measured on 2 B200s (32.8 MB per step and rank): 1 piece 0.094 ms per step, 2 pieces 0.099, 4: 0.107, 8: 0.107,
            // 16: 0.167 -- the enqueue cost and engine contention outweigh any gain; RG_GATHER_SPLIT keeps the experiment.
            int split = c->gather_split > 0 ? c->gather_split : 1;
            split = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)split, bytes >> 20));
            const uint64_t piece = (((bytes + split - 1) / split) + 255) & ~(uint64_t)255;
            int used = 1, next = 0;
            for (int q = 0; q < npeers; q++)
                for (uint64_t o = 0; o < bytes; o += piece) {
                    const int si = next++ % RG_PEER_STREAMS;
                    if (si >= used) { CU(c, cudaStreamWaitEvent(c->s_peer[si], c->ev_src, 0)); used = si + 1; }
                    CU(c, cudaMemcpyAsync((char*)peer_base[q] + dst_offset + o, (const char*)src + o, std::min<uint64_t>(piece, bytes - o),
                                          cudaMemcpyDeviceToDevice, c->s_peer[si]));
                }
            for (int si = 1; si < used; si++) {
                CU(c, cudaEventRecord(c->ev_peer_join[si], c->s_peer[si]));
                CU(c, cudaStreamWaitEvent(s0, c->ev_peer_join[si], 0));
            }
        } else {
            PeerPtrs pp;
            bool a16 = !(bytes & 15) && !(dst_offset & 15) && !((uintptr_t)src & 15);
            for (int q = 0; q < npeers; q++) { pp.p[q] = (double*)peer_base[q]; a16 = a16 && !((uintptr_t)peer_base[q] & 15); }
            const int grid = 2 * c->sm_count;
            if (a16) k_peer_scatter16<<<grid, 256, 0, s0>>>((const double2*)src, bytes / 16, pp, npeers, dst_offset / 16);
            else k_peer_scatter8<<<grid, 256, 0, s0>>>((const double*)src, bytes / 8, pp, npeers, dst_offset / 8);
            CU(c, cudaGetLastError());
            c->launches++;
        }
    }
    CU(c, cudaEventRecord(c->ev_gather[slot], s0));
    c->gather_pending[slot] = true;
    return RG_OK;
}

// Make `cuda_stream` (any stream of this device, e.g. the one a collective barrier is issued on) wait for the gather last issued on
// `slot`: cross-rank completion = every rank's outgoing copies done + a barrier, without touching the evaluation stream.
extern "C" int rg_gather_wait_on(rg_ctx* c, int32_t slot, void* cuda_stream) {
    if (!c || slot < 0 || slot > 1) return RG_ERR_INVALID;
    if (!c->gather_pending[slot]) return RG_OK;
    CU(c, cudaStreamWaitEvent((cudaStream_t)cuda_stream, c->ev_gather[slot], 0));
    return RG_OK;
}

extern "C" int rg_gather_wait(rg_ctx* c, int32_t slot) {
    if (!c || slot < 0 || slot > 1) return RG_ERR_INVALID;
    if (!c->gather_pending[slot]) return RG_OK;
    CU(c, cudaStreamWaitEvent(c->stream, c->ev_gather[slot], 0));
    return RG_OK;
}
