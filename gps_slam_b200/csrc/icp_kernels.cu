#include "icp.h"
namespace icp
{
struct Tracker { int kind; };
Tracker *create_tracker(int kind, int, int, float, float) { Tracker *t = new Tracker(); t->kind = kind; return t; }
void destroy_tracker(Tracker *t) { delete t; }
__global__ void k_convert_depth(const short *__restrict__ in, float *__restrict__ out, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { short d = in[i]; out[i] = d <= 0 ? -1.0f : (float)d * (1.0f / 1000.0f) + 0.0f; }
}
void convert_depth(const short *depth_mm, float *depth_f, int W, int H, cudaStream_t st)
{
    int n = W * H;
    k_convert_depth<<<(n + 255) / 256, 256, 0, st>>>(depth_mm, depth_f, n);
}
int track_camera(Tracker *, const float *, const float4 *, const float4 *, float, float, float, float, const Mat4 &, int, se3::Pose *, cudaStream_t)
{
    return gs_set_error(__FILE__, __LINE__, "ICP tracker not built yet");
}
} // namespace icp
