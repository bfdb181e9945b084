// Peer-memory exchange segment + cross-GPU barrier (see gs_comm.h).  C ABI: gsb_comm_* in include/gpsslam_b200.h.
#include <cstdio>
#include <cstring>
#include <new>

#include "../../include/gpsslam_b200.h"
#include "common.cuh"
#include "gs.h"
#include "gs_comm.h"

namespace gs
{

// Every lane q < world signals peer q ("rank r has reached barrier `epoch`") and waits for peer q's signal.  The spin is bounded
// (~20 s of SM clocks): a peer that never arrives turns into an error flag on the host, not into a hung GPU.
__global__ void k_comm_barrier(CommView c, unsigned epoch, int *err)
{
    const int q = threadIdx.x;
    if (q >= c.world)
        return;
    __threadfence_system();
    unsigned *remote = c.flags[q] + c.rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(epoch) : "memory");
    const unsigned *mine = c.flags[c.rank] + q;
    const long long t0 = clock64();
    for (;;)
    {
        unsigned v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
        if ((int)(v - epoch) >= 0)
            break;
        if (clock64() - t0 > 40000000000LL)
        {
            *err = 1;
            break;
        }
        __nanosleep(200);
    }
    __threadfence_system();
}

void comm_barrier(gsb_comm *c, cudaStream_t st)
{
    c->epoch++;
    GS_COUNT_LAUNCHES(1);
    k_comm_barrier<<<1, 32, 0, st>>>(c->view, c->epoch, c->errHost);
}

} // namespace gs

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

extern "C" int gsb_comm_create(int device, int rank, int world, int width, int height, gsb_comm_t **out)
{
    if (!out || world < 1 || world > gs::COMM_MAX_WORLD || rank < 0 || rank >= world || width <= 0 || height <= 0)
        return gs_set_error(__FILE__, __LINE__, "invalid argument (world <= 16)");
    GS_CUDA_OK(cudaSetDevice(device));
    gsb_comm *c = new (std::nothrow) gsb_comm();
    if (!c)
        return gs_set_error(__FILE__, __LINE__, "out of host memory");
    memset(c, 0, sizeof *c);
    c->device = device, c->rank = rank, c->world = world, c->W = width, c->H = height;
    const int tw = (width + gs::TILE - 1) / gs::TILE, th = (height + gs::TILE - 1) / gs::TILE;
    c->T = tw * th;
    const size_t P = (size_t)width * height;
    const size_t slotFloats = (size_t)c->T * 256 * 5;
    c->offGather = 4096; // flags live in the first page
    c->offVout = align_up(c->offGather + (size_t)world * slotFloats * sizeof(float), 4096);
    c->offLoss = align_up(c->offVout + 2 * P * sizeof(float4), 4096);
    c->segBytes = align_up(c->offLoss + (size_t)c->T * sizeof(float), 4096);
    if (cudaMalloc((void **)&c->seg, c->segBytes) != cudaSuccess)
    {
        delete c;
        return gs_set_error(__FILE__, __LINE__, "exchange segment allocation failed");
    }
    cudaMemset(c->seg, 0, c->segBytes);
    if (cudaHostAlloc((void **)&c->errHost, sizeof(int), cudaHostAllocMapped) != cudaSuccess)
    {
        cudaFree(c->seg);
        delete c;
        return gs_set_error(__FILE__, __LINE__, "pinned allocation failed");
    }
    *c->errHost = 0;
    if (cudaMalloc((void **)&c->viewDev, sizeof(gs::CommView)) != cudaSuccess)
    {
        cudaFree(c->seg);
        cudaFreeHost(c->errHost);
        delete c;
        return gs_set_error(__FILE__, __LINE__, "allocation failed");
    }
    c->peer[rank] = c->seg;
    c->attached = world == 1;
    gs::CommView &v = c->view;
    v.rank = rank, v.world = world;
    v.tilesPerRank = (c->T + world - 1) / world;
    v.slotFloats = slotFloats;
    if (world == 1)
    {
        v.gather[0] = (float *)(c->seg + c->offGather), v.vout[0] = (float4 *)(c->seg + c->offVout);
        v.lossTile[0] = (float *)(c->seg + c->offLoss), v.flags[0] = (unsigned *)c->seg;
        cudaMemcpy(c->viewDev, &v, sizeof v, cudaMemcpyHostToDevice);
    }
    GS_CUDA_OK(cudaDeviceSynchronize());
    *out = c;
    return 0;
}

static void fill_view(gsb_comm *c)
{
    for (int q = 0; q < c->world; q++)
    {
        c->view.gather[q] = (float *)(c->peer[q] + c->offGather);
        c->view.vout[q] = (float4 *)(c->peer[q] + c->offVout);
        c->view.lossTile[q] = (float *)(c->peer[q] + c->offLoss);
        c->view.flags[q] = (unsigned *)c->peer[q];
    }
    cudaMemcpy(c->viewDev, &c->view, sizeof c->view, cudaMemcpyHostToDevice);
    c->attached = true;
}

extern "C" int gsb_comm_export(gsb_comm_t *c, void *handle64)
{
    if (!c || !handle64)
        return gs_set_error(__FILE__, __LINE__, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    GS_CUDA_OK(cudaIpcGetMemHandle(&h, c->seg));
    memcpy(handle64, &h, 64);
    return 0;
}

extern "C" int gsb_comm_attach(gsb_comm_t *c, const void *handles)
{
    if (!c || !handles)
        return gs_set_error(__FILE__, __LINE__, "null argument");
    GS_CUDA_OK(cudaSetDevice(c->device));
    for (int q = 0; q < c->world; q++)
    {
        if (q == c->rank)
            continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)handles + (size_t)q * 64, 64);
        void *p = nullptr;
        GS_CUDA_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        c->peer[q] = (char *)p;
        c->ipcOpened[q] = true;
    }
    fill_view(c);
    return 0;
}

extern "C" int gsb_comm_attach_local(gsb_comm_t *c, gsb_comm_t *const *peers)
{
    if (!c || !peers)
        return gs_set_error(__FILE__, __LINE__, "null argument");
    for (int q = 0; q < c->world; q++)
    {
        if (!peers[q] || peers[q]->world != c->world || peers[q]->rank != q || peers[q]->segBytes != c->segBytes)
            return gs_set_error(__FILE__, __LINE__, "peer list does not match this communicator");
        if (peers[q]->device != c->device)
        {
            int can = 0;
            GS_CUDA_OK(cudaDeviceCanAccessPeer(&can, c->device, peers[q]->device));
            if (!can)
                return gs_set_error(__FILE__, __LINE__, "no peer access between the devices");
            GS_CUDA_OK(cudaSetDevice(c->device));
            cudaError_t e = cudaDeviceEnablePeerAccess(peers[q]->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                return gs_set_error(__FILE__, __LINE__, cudaGetErrorString(e));
            cudaGetLastError();
        }
        c->peer[q] = peers[q]->seg;
    }
    fill_view(c);
    return 0;
}

extern "C" int gsb_comm_barrier(gsb_comm_t *c, void *stream)
{
    if (!c || !c->attached)
        return gs_set_error(__FILE__, __LINE__, "communicator not attached");
    gs::comm_barrier(c, (cudaStream_t)stream);
    GS_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int gsb_comm_error(gsb_comm_t *c) { return c ? *c->errHost : 1; }

extern "C" void gsb_comm_destroy(gsb_comm_t *c)
{
    if (!c)
        return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (int q = 0; q < c->world; q++)
        if (c->ipcOpened[q])
            cudaIpcCloseMemHandle(c->peer[q]);
    cudaFree(c->seg);
    cudaFree(c->viewDev);
    cudaFreeHost(c->errHost);
    delete c;
}
