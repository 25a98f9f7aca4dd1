// kfusion::cuda::TsdfVolume::raycast (src/kfusion/tsdf_volume.cpp:95-129 -> src/kfusion/cuda/tsdf_volume.cu:126-386): the
// zero crossing of the TSDF along every pixel's ray, refined by trilinear interpolation, with the normal from central
// differences of the interpolated TSDF -- the step after integration (rendering the fused canonical model, and the
// model-side maps of KinFu's ICP).
//
// Canonical arithmetic (the reference is built with --ftz / --prec-div=false / --prec-sqrt=false and its dot products come
// from an un-vendored header, so no CPU can reproduce its bits): every operation of the reference's expressions rounded
// once, in the order written there; IEEE division and square root; normalized(v) = v / sqrt(dot(v, v)); dot(a, b) =
// (a.x b.x + a.y b.y) + a.z b.z.  One deviation that only removes undefined behaviour: fetch_tsdf clamps its voxel index
// into the volume (the reference reads out of bounds when rounding pushes the entry point of a ray just outside).
#include <math_constants.h>

#include "dfu_internal.h"
#include "dfu_math.cuh"

using namespace dfu;

namespace {

struct RayArgs {
    const uint32_t* vol;
    int dx, dy, dz;
    float vsx, vsy, vsz;
    float ivx, ivy, ivz;     // 1 / voxel size (tsdf_volume.cu:361)
    float bmx, bmy, bmz;     // box_max = volume_size - voxel_size (:219)
    float time_step;         // trunc * raycast_step_factor (:359)
    float gdx, gdy, gdz;     // gradient_delta = voxel * gradient_delta_factor (:360)
    float R[9], T[3];        // cam2vol
    float Ri[9];             // its inverse rotation
    float finvx, finvy, cx, cy;
    int rows, cols;
    float4* points; size_t ppitch;
    uint16_t* depth; size_t dpitch;
    float4* normals; size_t npitch;
};

DFU_DEV float dot3(float ax, float ay, float az, float bx, float by, float bz) { return fadd(fadd(fmul(ax, bx), fmul(ay, by)), fmul(az, bz)); }
DFU_DEV float tsdf_at(const RayArgs& a, int x, int y, int z) {
    return __half2float(__ushort_as_half((unsigned short) (__ldg(&a.vol[(size_t) x + (size_t) y * a.dx + (size_t) z * a.dx * a.dy]) & 0xffffu)));
}
// fetch_tsdf (:190-196): nearest voxel, round to nearest even
DFU_DEV float fetch_tsdf(const RayArgs& a, float px, float py, float pz) {
    const int x = min(max(__float2int_rn(fmul(px, a.ivx)), 0), a.dx - 1);
    const int y = min(max(__float2int_rn(fmul(py, a.ivy)), 0), a.dy - 1);
    const int z = min(max(__float2int_rn(fmul(pz, a.ivz)), 0), a.dz - 1);
    return tsdf_at(a, x, y, z);
}
// interpolate (:147-171): trilinear, NaN outside
DFU_DEV float interpolate(const RayArgs& a, float cx, float cy, float cz) {
    const int gx = __float2int_rd(cx), gy = __float2int_rd(cy), gz = __float2int_rd(cz);
    if (gx < 0 || gx >= a.dx - 1 || gy < 0 || gy >= a.dy - 1 || gz < 0 || gz >= a.dz - 1) return CUDART_NAN_F;
    const float fa = fsub(cx, (float) gx), fb = fsub(cy, (float) gy), fc = fsub(cz, (float) gz);
    const float na = fsub(1.f, fa), nb = fsub(1.f, fb), nc = fsub(1.f, fc);
    float t = 0.f;
    t = fadd(t, fmul(fmul(fmul(tsdf_at(a, gx, gy, gz), na), nb), nc));
    t = fadd(t, fmul(fmul(fmul(tsdf_at(a, gx, gy, gz + 1), na), nb), fc));
    t = fadd(t, fmul(fmul(fmul(tsdf_at(a, gx, gy + 1, gz), na), fb), nc));
    t = fadd(t, fmul(fmul(fmul(tsdf_at(a, gx, gy + 1, gz + 1), na), fb), fc));
    t = fadd(t, fmul(fmul(fmul(tsdf_at(a, gx + 1, gy, gz), fa), nb), nc));
    t = fadd(t, fmul(fmul(fmul(tsdf_at(a, gx + 1, gy, gz + 1), fa), nb), fc));
    t = fadd(t, fmul(fmul(fmul(tsdf_at(a, gx + 1, gy + 1, gz), fa), fb), nc));
    t = fadd(t, fmul(fmul(fmul(tsdf_at(a, gx + 1, gy + 1, gz + 1), fa), fb), fc));
    return t;
}
DFU_DEV float interp_m(const RayArgs& a, float px, float py, float pz) { return interpolate(a, fmul(px, a.ivx), fmul(py, a.ivy), fmul(pz, a.ivz)); }

// Measured alternatives (round 2, 640x480 rays into the 512^3 bench volume, ncu gpu__time_duration): this plain march 143 us;
// six steps laid out ahead with their fetches in flight and 8x4-pixel warp patches 253 us (96 registers, wasted fetches past the
// crossing); two steps ahead on 32x1 rows 197 us.  The march is bound by the 32-byte sectors its 4-byte nearest-voxel samples
// pull through L1/L2 (three quarters of the rays cross the whole volume through empty space), not by the latency of one ray,
// so look-ahead does not help; skipping empty space needs an occupancy structure kept by the integrator.
__global__ void __launch_bounds__(256) raycast_kernel(const RayArgs a) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= a.cols || y >= a.rows) return;
    const float qnan = CUDART_NAN_F;
    float4 P = make_float4(qnan, qnan, qnan, qnan), Nn = P;
    uint16_t D = 0;
    // ray_dir = normalized(aff.R * reproj(x, y, 1.f)) (:213)
    const float ux = fmul(fmul(1.f, fsub((float) x, a.cx)), a.finvx), uy = fmul(fmul(1.f, fsub((float) y, a.cy)), a.finvy), uz = 1.f;
    float rdx = dot3(a.R[0], a.R[1], a.R[2], ux, uy, uz), rdy = dot3(a.R[3], a.R[4], a.R[5], ux, uy, uz),
          rdz = dot3(a.R[6], a.R[7], a.R[8], ux, uy, uz);
    const float len = __fsqrt_rn(dot3(rdx, rdy, rdz, rdx, rdy, rdz));
    rdx = __fdiv_rn(rdx, len); rdy = __fdiv_rn(rdy, len); rdz = __fdiv_rn(rdz, len);
    const float ox = a.T[0], oy = a.T[1], oz = a.T[2];
    // intersect (:126-144)
    const float ix = __fdiv_rn(1.f, rdx), iy = __fdiv_rn(1.f, rdy), iz = __fdiv_rn(1.f, rdz);
    const float tbx = fmul(ix, fsub(0.f, ox)), tby = fmul(iy, fsub(0.f, oy)), tbz = fmul(iz, fsub(0.f, oz));
    const float ttx = fmul(ix, fsub(a.bmx, ox)), tty = fmul(iy, fsub(a.bmy, oy)), ttz = fmul(iz, fsub(a.bmz, oz));
    const float mnx = fminf(ttx, tbx), mny = fminf(tty, tby), mnz = fminf(ttz, tbz);
    const float mxx = fmaxf(ttx, tbx), mxy = fmaxf(tty, tby), mxz = fmaxf(ttz, tbz);
    float tmin = fmaxf(fmaxf(mnx, mny), fmaxf(mnx, mnz));
    float tmax = fminf(fminf(mxx, mxy), fminf(mxx, mxz));
    tmin = fmaxf(0.f, tmin);
    if (!(tmin >= tmax)) {  // `if (tmin >= tmax) return;` (:227)
        tmax = fsub(tmax, a.time_step);
        const float vsx_ = fmul(rdx, a.time_step), vsy_ = fmul(rdy, a.time_step), vsz_ = fmul(rdz, a.time_step);
        float nx = fadd(ox, fmul(rdx, tmin)), ny = fadd(oy, fmul(rdy, tmin)), nz = fadd(oz, fmul(rdz, tmin));
        float tsdf_next = fetch_tsdf(a, nx, ny, nz);
        for (float tcurr = tmin; tcurr < tmax; tcurr = fadd(tcurr, a.time_step)) {
            const float tsdf_curr = tsdf_next;
            const float cxm = nx, cym = ny, czm = nz;
            nx = fadd(nx, vsx_); ny = fadd(ny, vsy_); nz = fadd(nz, vsz_);
            tsdf_next = fetch_tsdf(a, nx, ny, nz);
            if (tsdf_curr < 0.f && tsdf_next > 0.f) break;
            if (tsdf_curr > 0.f && tsdf_next < 0.f) {
                const float Ft = interp_m(a, cxm, cym, czm), Ftdt = interp_m(a, nx, ny, nz);
                const float Ts = fsub(tcurr, __fdiv_rn(fmul(a.time_step, Ft), fsub(Ftdt, Ft)));
                const float vx = fadd(ox, fmul(rdx, Ts)), vy = fadd(oy, fmul(rdy, Ts)), vz = fadd(oz, fmul(rdz, Ts));
                // compute_normal (:307-324)
                float gx = __fdiv_rn(fsub(interp_m(a, fadd(vx, a.gdx), vy, vz), interp_m(a, fsub(vx, a.gdx), vy, vz)), a.gdx);
                float gy = __fdiv_rn(fsub(interp_m(a, vx, fadd(vy, a.gdy), vz), interp_m(a, vx, fsub(vy, a.gdy), vz)), a.gdy);
                float gz = __fdiv_rn(fsub(interp_m(a, vx, vy, fadd(vz, a.gdz)), interp_m(a, vx, vy, fsub(vz, a.gdz))), a.gdz);
                const float gl = __fsqrt_rn(dot3(gx, gy, gz, gx, gy, gz));
                gx = __fdiv_rn(gx, gl); gy = __fdiv_rn(gy, gl); gz = __fdiv_rn(gz, gl);
                const float prod = fmul(fmul(gx, gy), gz);
                if (prod == prod) {  // !isnan
                    const float dxv = fsub(vx, ox), dyv = fsub(vy, oy), dzv = fsub(vz, oz);
                    Nn = make_float4(dot3(a.Ri[0], a.Ri[1], a.Ri[2], gx, gy, gz), dot3(a.Ri[3], a.Ri[4], a.Ri[5], gx, gy, gz),
                                     dot3(a.Ri[6], a.Ri[7], a.Ri[8], gx, gy, gz), 0.f);
                    P = make_float4(dot3(a.Ri[0], a.Ri[1], a.Ri[2], dxv, dyv, dzv), dot3(a.Ri[3], a.Ri[4], a.Ri[5], dxv, dyv, dzv),
                                    dot3(a.Ri[6], a.Ri[7], a.Ri[8], dxv, dyv, dzv), 0.f);
                    const float mm = fmul(P.z, 1000.f);  // static_cast<ushort>(vertex.z * 1000), saturated
                    D = (uint16_t) min(max((int) mm, 0), 65535);
                }
                break;
            }
        }
    }
    if (a.points) *reinterpret_cast<float4*>(reinterpret_cast<char*>(a.points) + (size_t) y * a.ppitch + sizeof(float4) * x) = P;
    if (a.depth) *reinterpret_cast<uint16_t*>(reinterpret_cast<char*>(a.depth) + (size_t) y * a.dpitch + sizeof(uint16_t) * x) = D;
    *reinterpret_cast<float4*>(reinterpret_cast<char*>(a.normals) + (size_t) y * a.npitch + sizeof(float4) * x) = Nn;
}

}  // namespace

extern "C" int dfu_tsdf_raycast(const void* volume, const int dims_host[3], const float voxel_size_host[3], float trunc_dist,
                                const float cam2vol_host[12], const float rinv_host[9], const float intr_host[4], int rows, int cols,
                                float raycast_step_factor, float gradient_delta_factor, float* points4, size_t points_pitch_bytes,
                                uint16_t* depth, size_t depth_pitch_bytes, float* normals4, size_t normals_pitch_bytes,
                                dfu_stream stream) {
    DFU_REQUIRE(volume && dims_host && voxel_size_host && cam2vol_host && rinv_host && intr_host && normals4, DFU_ERR_INVALID, "NULL argument");
    DFU_REQUIRE(points4 || depth, DFU_ERR_INVALID, "need a points image, a depth image, or both");
    DFU_REQUIRE(rows > 0 && cols > 0 && dims_host[0] > 1 && dims_host[1] > 1 && dims_host[2] > 1, DFU_ERR_INVALID, "bad size");
    DFU_REQUIRE(raycast_step_factor > 0.f && gradient_delta_factor > 0.f && trunc_dist > 0.f, DFU_ERR_INVALID, "bad step / delta / trunc");
    DFU_GUARD(dfu_device_of(volume));
    RayArgs a{};
    a.vol = static_cast<const uint32_t*>(volume);
    a.dx = dims_host[0]; a.dy = dims_host[1]; a.dz = dims_host[2];
    a.vsx = voxel_size_host[0]; a.vsy = voxel_size_host[1]; a.vsz = voxel_size_host[2];
    a.ivx = 1.f / a.vsx; a.ivy = 1.f / a.vsy; a.ivz = 1.f / a.vsz;
    a.bmx = a.vsx * (float) a.dx - a.vsx; a.bmy = a.vsy * (float) a.dy - a.vsy; a.bmz = a.vsz * (float) a.dz - a.vsz;
    a.time_step = trunc_dist * raycast_step_factor;
    a.gdx = a.vsx * gradient_delta_factor; a.gdy = a.vsy * gradient_delta_factor; a.gdz = a.vsz * gradient_delta_factor;
    for (int i = 0; i < 9; ++i) {
        a.R[i] = cam2vol_host[i];
        a.Ri[i] = rinv_host[i];
    }
    for (int i = 0; i < 3; ++i) a.T[i] = cam2vol_host[9 + i];
    a.finvx = 1.f / intr_host[0]; a.finvy = 1.f / intr_host[1]; a.cx = intr_host[2]; a.cy = intr_host[3];
    a.rows = rows; a.cols = cols;
    a.points = reinterpret_cast<float4*>(points4); a.ppitch = points_pitch_bytes;
    a.depth = depth; a.dpitch = depth_pitch_bytes;
    a.normals = reinterpret_cast<float4*>(normals4); a.npitch = normals_pitch_bytes;
    dim3 block(32, 8), grid(div_up(cols, 32), div_up(rows, 8));
    raycast_kernel<<<grid, block, 0, as_stream(stream)>>>(a);
    DFU_LAUNCH_OK();
    return DFU_OK;
}
