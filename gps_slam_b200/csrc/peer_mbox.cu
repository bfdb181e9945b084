// Peer mailbox: a byte segment per rank that every other rank of the box can write into over NVLink, plus counters for hand-shakes.
// Used by the functional split of the SLAM loop: the rank that owns the TSDF side stores the camera maps of a cycle
// into the Gaussian ranks' mailboxes and raises a counter; the Gaussian ranks' streams wait for the counter, and raise one of the TSDF
// rank's counters when they are done with a set of maps.  C ABI: gsb_mbox_* in include/gpsslam_b200.h.  Same mapping pattern as
// gsb_comm_* (CUDA IPC between processes, plain pointers between engines of one process).
#include <cstring>
#include <new>

#include "../../include/gpsslam_b200.h"
#include "common.cuh"

namespace
{
constexpr int MBOX_MAX_WORLD = 16;
constexpr int MBOX_FLAGS = 256;          // counters per rank (first page of the segment)
constexpr size_t MBOX_DATA_OFF = 4096;

__global__ void k_mbox_signal(unsigned *remote, unsigned value)
{
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(value) : "memory");
}

// the stream does not proceed until the local counter has reached `value` (bounded spin: a peer that never signals becomes an error flag)
__global__ void k_mbox_wait(const unsigned *mine, unsigned value, int *err)
{
    const long long t0 = clock64();
    for (;;)
    {
        unsigned v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
        if ((int)(v - value) >= 0)
            break;
        if (clock64() - t0 > 120000000000LL)
        {
            *err = 1;
            break;
        }
        __nanosleep(500);
    }
    __threadfence_system();
}
} // namespace

struct gsb_mbox
{
    int device, rank, world;
    size_t bytes, segBytes;
    char *seg;
    char *peer[MBOX_MAX_WORLD];
    bool ipcOpened[MBOX_MAX_WORLD];
    bool attached;
    int *errHost;
};

extern "C" int gsb_mbox_create(int device, int rank, int world, size_t bytes, gsb_mbox_t **out)
{
    if (!out || world < 1 || world > MBOX_MAX_WORLD || rank < 0 || rank >= world)
        return gs_set_error(__FILE__, __LINE__, "invalid argument (world <= 16)");
    GS_CUDA_OK(cudaSetDevice(device));
    gsb_mbox *m = new (std::nothrow) gsb_mbox();
    if (!m)
        return gs_set_error(__FILE__, __LINE__, "out of host memory");
    memset(m, 0, sizeof *m);
    m->device = device, m->rank = rank, m->world = world, m->bytes = bytes;
    m->segBytes = MBOX_DATA_OFF + (bytes + 4095) / 4096 * 4096;
    if (cudaMalloc((void **)&m->seg, m->segBytes) != cudaSuccess)
    {
        delete m;
        return gs_set_error(__FILE__, __LINE__, "mailbox allocation failed");
    }
    cudaMemset(m->seg, 0, MBOX_DATA_OFF);
    if (cudaHostAlloc((void **)&m->errHost, sizeof(int), cudaHostAllocMapped) != cudaSuccess)
    {
        cudaFree(m->seg);
        delete m;
        return gs_set_error(__FILE__, __LINE__, "pinned allocation failed");
    }
    *m->errHost = 0;
    m->peer[rank] = m->seg;
    m->attached = world == 1;
    GS_CUDA_OK(cudaDeviceSynchronize());
    *out = m;
    return 0;
}

extern "C" int gsb_mbox_export(gsb_mbox_t *m, void *handle64)
{
    if (!m || !handle64)
        return gs_set_error(__FILE__, __LINE__, "null argument");
    cudaIpcMemHandle_t h;
    GS_CUDA_OK(cudaIpcGetMemHandle(&h, m->seg));
    memcpy(handle64, &h, 64);
    return 0;
}

extern "C" int gsb_mbox_attach(gsb_mbox_t *m, const void *handles)
{
    if (!m || !handles)
        return gs_set_error(__FILE__, __LINE__, "null argument");
    GS_CUDA_OK(cudaSetDevice(m->device));
    for (int q = 0; q < m->world; q++)
    {
        if (q == m->rank || m->ipcOpened[q])
            continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)handles + (size_t)q * 64, 64);
        void *p = nullptr;
        GS_CUDA_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        m->peer[q] = (char *)p;
        m->ipcOpened[q] = true;
    }
    m->attached = true;
    return 0;
}

extern "C" int gsb_mbox_attach_local(gsb_mbox_t *m, gsb_mbox_t *const *peers)
{
    if (!m || !peers)
        return gs_set_error(__FILE__, __LINE__, "null argument");
    for (int q = 0; q < m->world; q++)
    {
        if (!peers[q] || peers[q]->world != m->world || peers[q]->rank != q)
            return gs_set_error(__FILE__, __LINE__, "peer list does not match this mailbox");
        if (peers[q]->device != m->device)
        {
            int can = 0;
            GS_CUDA_OK(cudaDeviceCanAccessPeer(&can, m->device, peers[q]->device));
            if (!can)
                return gs_set_error(__FILE__, __LINE__, "no peer access between the devices");
            GS_CUDA_OK(cudaSetDevice(m->device));
            cudaError_t e = cudaDeviceEnablePeerAccess(peers[q]->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                return gs_set_error(__FILE__, __LINE__, cudaGetErrorString(e));
            cudaGetLastError();
        }
        m->peer[q] = peers[q]->seg;
    }
    m->attached = true;
    return 0;
}

extern "C" void *gsb_mbox_local(gsb_mbox_t *m) { return m ? (void *)(m->seg + MBOX_DATA_OFF) : nullptr; }

extern "C" int gsb_mbox_put(gsb_mbox_t *m, int dst_rank, size_t dst_offset, const void *src_dev, size_t bytes, void *stream)
{
    if (!m || !m->attached || dst_rank < 0 || dst_rank >= m->world || !src_dev)
        return gs_set_error(__FILE__, __LINE__, "mailbox not attached / bad destination");
    if (dst_offset + bytes > m->bytes)
        return gs_set_error(__FILE__, __LINE__, "put beyond the destination mailbox");
    GS_CUDA_OK(cudaSetDevice(m->device));
    GS_CUDA_OK(cudaMemcpyAsync(m->peer[dst_rank] + MBOX_DATA_OFF + dst_offset, src_dev, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return 0;
}

extern "C" int gsb_mbox_signal(gsb_mbox_t *m, int dst_rank, int flag, unsigned value, void *stream)
{
    if (!m || !m->attached || dst_rank < 0 || dst_rank >= m->world || flag < 0 || flag >= MBOX_FLAGS)
        return gs_set_error(__FILE__, __LINE__, "mailbox not attached / bad flag");
    GS_CUDA_OK(cudaSetDevice(m->device));
    GS_COUNT_LAUNCHES(1);
    k_mbox_signal<<<1, 1, 0, (cudaStream_t)stream>>>(reinterpret_cast<unsigned *>(m->peer[dst_rank]) + flag, value);
    GS_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int gsb_mbox_wait(gsb_mbox_t *m, int flag, unsigned value, void *stream)
{
    if (!m || !m->attached || flag < 0 || flag >= MBOX_FLAGS)
        return gs_set_error(__FILE__, __LINE__, "mailbox not attached / bad flag");
    GS_CUDA_OK(cudaSetDevice(m->device));
    GS_COUNT_LAUNCHES(1);
    k_mbox_wait<<<1, 1, 0, (cudaStream_t)stream>>>(reinterpret_cast<const unsigned *>(m->seg) + flag, value, m->errHost);
    GS_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int gsb_mbox_error(gsb_mbox_t *m) { return m ? *m->errHost : 1; }

extern "C" void gsb_mbox_destroy(gsb_mbox_t *m)
{
    if (!m)
        return;
    cudaSetDevice(m->device);
    cudaDeviceSynchronize();
    for (int q = 0; q < m->world; q++)
        if (m->ipcOpened[q])
            cudaIpcCloseMemHandle(m->peer[q]);
    cudaFree(m->seg);
    cudaFreeHost(m->errHost);
    delete m;
}
