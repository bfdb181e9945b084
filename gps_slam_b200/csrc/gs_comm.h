// Peer-memory exchange of the multi-GPU Gaussian path: one process per GPU, every rank maps every other rank's
// exchange segment (CUDA IPC over NVLink / NVSwitch), kernels store straight into peer memory.
//
// Per optimiser iteration, G ranks:
//   1. k_raster_fwd<PUSH>  : each rank rasterises ITS Gaussians over the whole image; the epilogue stores every tile's partial sums
//                            (rgb, depth, weight) into slot `rank` of the gather region of the tile's OWNER rank (tiles are dealt to
//                            ranks in contiguous ranges) -- the reduce-scatter is the rasteriser's own output write;
//   2. k_comm_barrier      : flag exchange through peer memory (release / acquire at system scope, bounded spin);
//   3. k_composite_x       : the owner sums the G slots of its tiles in rank order (every rank therefore gets bit-identical values),
//                            composites with the TSDF render, takes the L1 loss and dL/d(render) and stores the 16-byte gradient record
//                            and the tile loss into EVERY rank's v_out / lossTile (the all-gather);
//   4. k_comm_barrier, then backward + Adam run locally on each rank's own Gaussians.
// A render (spawn mask, evaluation) pushes every tile to every rank instead and each rank composites the whole image.
#pragma once
#include <cuda_runtime.h>

namespace gs
{
constexpr int COMM_MAX_WORLD = 16;

// what the kernels see: peer base pointers of the three regions, indexed by rank
struct CommView
{
    int rank, world;
    int tilesPerRank;                 // owner(tile) = min(world - 1, tile / tilesPerRank)
    size_t slotFloats;                // floats per gather slot = T * 256 * 5 (tile-major image: rgbd as float4, then the weights)
    float *gather[COMM_MAX_WORLD];    // [world slots][slotFloats]
    float4 *vout[COMM_MAX_WORLD];     // [2 P]
    float *lossTile[COMM_MAX_WORLD];  // [T]
    unsigned *flags[COMM_MAX_WORLD];  // [COMM_MAX_WORLD] arrival epochs, one per peer
};
__host__ __device__ inline int comm_owner(const CommView &c, int tile)
{
    int o = tile / c.tilesPerRank;
    return o < c.world ? o : c.world - 1;
}
} // namespace gs

struct gsb_comm
{
    int device, rank, world, W, H, T;
    size_t segBytes, offGather, offVout, offLoss;
    char *seg;                           // local segment (cudaMalloc)
    char *peer[gs::COMM_MAX_WORLD];      // mapped base of every rank's segment (own entry = seg)
    bool ipcOpened[gs::COMM_MAX_WORLD];
    bool attached;
    unsigned epoch;                      // barriers issued so far (identical on every rank: SPMD call sequence)
    int *errHost;                        // pinned: set by a barrier that timed out
    gs::CommView view;
    gs::CommView *viewDev;               // the same in device memory (kernels that take it by pointer)
};

namespace gs
{
void comm_barrier(gsb_comm *c, cudaStream_t st);
}
