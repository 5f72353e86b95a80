// raygen.cu — the step right before traversal: ray generation on the device.
//   k_gen_primary  src/bin/bvh_cpu.rs:72-84 (CPU twin of src/bin/bvh_trace.wgsl:224-233): one ray per pixel from
//                  camera.clip_to_world.
//   k_gen_shadow   src/bin/raytraced_shadows.wgsl:93-99: ray_new(pos + nor * 0.0001, light.position - pos) from
//                  G-buffer world positions / normals; direction NOT normalised.
//   k_gen_area_shadow  same origin, direction toward p0 + (p1-p0)*u + (p3-p0)*v on a rect area light
//                  (crates/pools/src/light.rs:28-52 corner order); the README's "raytraced shadows" TODO for area
//                  lights.  (u,v) are inputs so the sampling pattern stays the caller's.
//   k_rotate_z     shaders/compute_update.wgsl:12-27: transform = from_rotation_z(+-angle) * transform per listed
//                  instance (the per-frame animation that makes a TLAS rebuild necessary).
// Compiled with -fmad=false; glam semantics: Mat4*Vec4 = ((X*x + Y*y) + Z*z) + W*w, Vec3/f32 = 3 divisions,
// normalize = v * (1 / sqrt((x*x + y*y) + z*z)).
#include "common.cuh"

namespace {

struct Mat4 { float m[16]; };  // column-major

__device__ __forceinline__ void mat_vec(const Mat4& M, float x, float y, float z, float w, float* o) {
#pragma unroll
    for (int r = 0; r < 4; ++r) o[r] = ((M.m[r] * x + M.m[4 + r] * y) + M.m[8 + r] * z) + M.m[12 + r] * w;
}

__global__ void __launch_bounds__(256) k_gen_primary(Mat4 clip_to_world, uint32_t width, uint32_t height, float* ro, float* rd) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)width * height) return;
    // bvh_cpu.rs:74-76 — note `i / HEIGHT` (sic); identical to i / WIDTH for the reference's square 640 x 640 image
    float x = __fdiv_rn((float)(i % width), (float)width);
    float y = __fdiv_rn((float)(i / height), (float)height);
    x = (x - 0.5f) * 2.0f;
    y = (y - 0.5f) * -2.0f;
    float vp[4], vt[4];
    mat_vec(clip_to_world, x, y, 1.0f, 1.0f, vp);
    mat_vec(clip_to_world, x, y, 0.0f, 1.0f, vt);
    ro[3 * i] = __fdiv_rn(vp[0], vp[3]);
    ro[3 * i + 1] = __fdiv_rn(vp[1], vp[3]);
    ro[3 * i + 2] = __fdiv_rn(vp[2], vp[3]);
    const float len = __fsqrt_rn((vt[0] * vt[0] + vt[1] * vt[1]) + vt[2] * vt[2]);
    const float rl = __fdiv_rn(1.0f, len);
    rd[3 * i] = vt[0] * rl;
    rd[3 * i + 1] = vt[1] * rl;
    rd[3 * i + 2] = vt[2] * rl;
}

__global__ void __launch_bounds__(256) k_gen_shadow(const float* __restrict__ pos, const float* __restrict__ nor, size_t n,
                                                    float lx, float ly, float lz, float* ro, float* rd) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float p[3] = {pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]};
    const float l[3] = {lx, ly, lz};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        ro[3 * i + k] = p[k] + nor[3 * i + k] * 0.0001f;
        rd[3 * i + k] = l[k] - p[k];
    }
}

struct Rect { float p[4][3]; };

__global__ void __launch_bounds__(256) k_gen_area_shadow(const float* __restrict__ pos, const float* __restrict__ nor,
                                                         const float* __restrict__ uv, size_t n, Rect rc, float* ro, float* rd) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float u = uv[2 * i], v = uv[2 * i + 1];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float p = pos[3 * i + k];
        const float o = p + nor[3 * i + k] * 0.0001f;
        const float t = (rc.p[0][k] + (rc.p[1][k] - rc.p[0][k]) * u) + (rc.p[3][k] - rc.p[0][k]) * v;
        ro[3 * i + k] = o;
        rd[3 * i + k] = t - o;
    }
}

// new[j] = A * B[j] with WGSL's mat*vec as the four-term column sum, left to right
__device__ __forceinline__ void mat_mul44(const float* A, const float* B, float* out) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int r = 0; r < 4; ++r)
            out[4 * j + r] = ((A[r] * B[4 * j] + A[4 + r] * B[4 * j + 1]) + A[8 + r] * B[4 * j + 2]) + A[12 + r] * B[4 * j + 3];
}

__global__ void __launch_bounds__(128) k_rotate_z(Instance* inst, const uint32_t* ids, size_t n, float s, float c, int update_inverse) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Instance* in = inst + (ids ? ids[i] : (uint32_t)i);
    float T[16], R[16], out[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) T[k] = in->transform[k];
    // compute_update.wgsl:20-25: speed *= 1.0 when transform[3][2] > -15.0, else speed *= -1.0
    const float sg = (T[14] > -15.0f) ? s : -s;
#pragma unroll
    for (int k = 0; k < 16; ++k) R[k] = 0.0f;
    R[0] = c; R[1] = sg; R[4] = -sg; R[5] = c; R[10] = 1.0f; R[15] = 1.0f;  // math.wgsl from_rotation_z, column-major
    mat_mul44(R, T, out);
#pragma unroll
    for (int k = 0; k < 16; ++k) in->transform[k] = out[k];
    if (update_inverse) {
        float I[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) I[k] = in->inv_transform[k];
        R[1] = -sg; R[4] = sg;  // from_rotation_z(-angle)
        mat_mul44(I, R, out);
#pragma unroll
        for (int k = 0; k < 16; ++k) in->inv_transform[k] = out[k];
    }
}

}  // namespace

extern "C" {

int bvh_cuda_gen_primary_rays_dev(bvh_cuda_ctx* ctx, const float* clip_to_world, uint32_t width, uint32_t height, float* d_ray_o,
                                  float* d_ray_d, void* stream) {
    if (!ctx) return BVH_CUDA_EINVAL;
    if (!clip_to_world || !d_ray_o || !d_ray_d || width == 0 || height == 0) return ctx_fail(ctx, BVH_CUDA_EINVAL, "gen_primary_rays: bad argument");
    Mat4 M;
    for (int k = 0; k < 16; ++k) M.m[k] = clip_to_world[k];
    const size_t n = (size_t)width * height;
    DeviceGuard g(ctx->device);
    k_gen_primary<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(M, width, height, d_ray_o, d_ray_d);
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    return BVH_CUDA_OK;
}

int bvh_cuda_gen_shadow_rays_dev(bvh_cuda_ctx* ctx, const float* d_pos, const float* d_nor, size_t n, const float* light_pos,
                                 float* d_ray_o, float* d_ray_d, void* stream) {
    if (!ctx) return BVH_CUDA_EINVAL;
    if (!d_pos || !d_nor || !light_pos || !d_ray_o || !d_ray_d) return ctx_fail(ctx, BVH_CUDA_EINVAL, "gen_shadow_rays: null pointer");
    if (n == 0) return BVH_CUDA_OK;
    DeviceGuard g(ctx->device);
    k_gen_shadow<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_pos, d_nor, n, light_pos[0], light_pos[1], light_pos[2], d_ray_o, d_ray_d);
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    return BVH_CUDA_OK;
}

int bvh_cuda_gen_area_shadow_rays_dev(bvh_cuda_ctx* ctx, const float* d_pos, const float* d_nor, const float* d_uv, size_t n,
                                      const float* corners, float* d_ray_o, float* d_ray_d, void* stream) {
    if (!ctx) return BVH_CUDA_EINVAL;
    if (!d_pos || !d_nor || !d_uv || !corners || !d_ray_o || !d_ray_d) return ctx_fail(ctx, BVH_CUDA_EINVAL, "gen_area_shadow_rays: null pointer");
    if (n == 0) return BVH_CUDA_OK;
    Rect rc;
    for (int c = 0; c < 4; ++c)
        for (int k = 0; k < 3; ++k) rc.p[c][k] = corners[3 * c + k];
    DeviceGuard g(ctx->device);
    k_gen_area_shadow<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_pos, d_nor, d_uv, n, rc, d_ray_o, d_ray_d);
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    return BVH_CUDA_OK;
}

int bvh_cuda_instances_rotate_z_dev(bvh_cuda_ctx* ctx, Instance* d_instances, const uint32_t* d_ids, size_t n, float sin_a,
                                    float cos_a, int update_inverse, void* stream) {
    if (!ctx) return BVH_CUDA_EINVAL;
    if (!d_instances) return ctx_fail(ctx, BVH_CUDA_EINVAL, "instances_rotate_z: null pointer");
    if (n == 0) return BVH_CUDA_OK;
    if (n > 0xFFFFFFFFull) return ctx_fail(ctx, BVH_CUDA_EINVAL, "instances_rotate_z: too many instances");
    DeviceGuard g(ctx->device);
    k_rotate_z<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(d_instances, d_ids, n, sin_a, cos_a, update_inverse);
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    return BVH_CUDA_OK;
}

}  // extern "C"
