// Shared device/host types for the B200 SLAM engine.
//
// Layouts are wire-compatible with the reference's scene objects so that dumps and parity checks line up:
//   HashEntry  <-> ITMHashEntry      (InfiniTAM/ITMLib/Objects/Scene/ITMVoxelBlockHash.h:36-48)   16 B
//   Voxel      <-> ITMVoxel_s_rgb    (InfiniTAM/ITMLib/Objects/Scene/ITMVoxelTypes.h:41-69)        8 B
//   Mat4       <-> ORUtils::Matrix4  (InfiniTAM/ORUtils/Matrix.h:26-36), column-major m[col*4+row]
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define SDF_BLOCK_SIZE 8
#define SDF_BLOCK_SIZE3 512
#define SDF_BUCKET_NUM 0x100000
#define SDF_HASH_MASK 0xfffff
#define SDF_EXCESS_LIST_SIZE 0x20000
#define SDF_TOTAL_ENTRIES (SDF_BUCKET_NUM + SDF_EXCESS_LIST_SIZE)
#define SDF_DEFAULT_BLOCK_NUM 0x40000

struct __align__(16) HashEntry
{
    short px, py, pz, pad;
    int offset;
    int ptr;
};
static_assert(sizeof(HashEntry) == 16, "hash entry must be 16 B");

// 8-byte voxel: sdf(short) | w_depth(u8) | r g b (u8) | w_color(u8) | pad
struct __align__(8) Voxel
{
    short sdf;
    unsigned char w_depth;
    unsigned char r, g, b;
    unsigned char w_color;
    unsigned char pad;
};
static_assert(sizeof(Voxel) == 8, "voxel must be 8 B");

struct Mat4
{
    float m[16];
};

#ifdef __CUDACC__
#define GS_HD __host__ __device__ __forceinline__
#define GS_D __device__ __forceinline__

// r = M * (x,y,z,w) with the reference's summation order (Matrix.h:130-137): ((m0*x + m4*y) + m8*z) + m12*w
GS_HD float3 mat4_mul_point(const Mat4 &M, float x, float y, float z, float w)
{
    float3 r;
    r.x = M.m[0] * x + M.m[4] * y + M.m[8] * z + M.m[12] * w;
    r.y = M.m[1] * x + M.m[5] * y + M.m[9] * z + M.m[13] * w;
    r.z = M.m[2] * x + M.m[6] * y + M.m[10] * z + M.m[14] * w;
    return r;
}

// hashIndex (ITMRepresentationAccess.h:7-11)
GS_HD int hash_index(int bx, int by, int bz)
{
    return (int)((((unsigned)bx * 73856093u) ^ ((unsigned)by * 19349669u) ^ ((unsigned)bz * 83492791u)) & (unsigned)SDF_HASH_MASK);
}

GS_D HashEntry load_entry(const HashEntry *table, int idx)
{
    int4 v = __ldg(reinterpret_cast<const int4 *>(table) + idx);
    HashEntry e;
    e.px = (short)(v.x & 0xffff);
    e.py = (short)((unsigned)v.x >> 16);
    e.pz = (short)(v.y & 0xffff);
    e.pad = 0;
    e.offset = v.z;
    e.ptr = v.w;
    return e;
}
// same but through the coherent path (table may have been written earlier in the same kernel)
GS_D HashEntry load_entry_cg(const HashEntry *table, int idx)
{
    int4 v = __ldcg(reinterpret_cast<const int4 *>(table) + idx);
    HashEntry e;
    e.px = (short)(v.x & 0xffff);
    e.py = (short)((unsigned)v.x >> 16);
    e.pz = (short)(v.y & 0xffff);
    e.pad = 0;
    e.offset = v.z;
    e.ptr = v.w;
    return e;
}
#endif

#define GS_CUDA_OK(call)                                                                  \
    do                                                                                    \
    {                                                                                     \
        cudaError_t _e = (call);                                                          \
        if (_e != cudaSuccess)                                                            \
            return gs_set_error(__FILE__, __LINE__, cudaGetErrorString(_e));              \
    } while (0)

int gs_set_error(const char *file, int line, const char *msg);

// number of kernels this library has launched (read through gsb_launch_count(); bench.py reports it as gpu_launches)
extern long long g_gsb_launches;
#define GS_COUNT_LAUNCHES(n) (g_gsb_launches += (n))
