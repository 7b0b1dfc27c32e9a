// Batched, host-free ends of the generation loop (reference run_nerf.py:27-147 per-image loop,
// core/utils/ray_utils.py:83-136 kp_to_valid_rays, core/utils/skeleton_utils.py:700-787 cylinder_to_box_2d) and the
// backward of the forward-kinematics chain (core/utils/skeleton_utils.py:379-463, core/pose_opt.py:372-445):
//   pgn_cyl_bbox_kernel              cylinder (cx, cz, R, top, bot) -> integer pixel box, one thread per pose (fp64 like numpy)
//   pgn_generate_rays_batch_kernel   rays of B bboxes into ONE [N,11] batch + pose_idx[N] (one launch for the batch)
//   pgn_compose_frames_batch_kernel  B white-background frames from the batch's rgb/acc rows
//   pgn_pose_fk_backward_kernel      dL/d skts [B,24,4,4] -> dL/d bones [B,24,3]
// All HBM-bound streaming kernels (44 + 4 B written per ray; 12 B per pixel), grids sized in multiples of the SM count.
#include "pgn_common.cuh"
#include "pgn_kernels.h"

namespace {

__constant__ int c_parents[PGN_J] = {0, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21};

// cylinder_to_box_2d: 50 points per cap (np.linspace(0, 2 pi, 50)), projected with w2c and K = diag(f, f, 1), floor / ceil,
// + (int(W/2), int(H/2)), clipped to [0, W-1] x [0, H-1].  w2c: row-major 4x4 (host-computed inverse of the swapped c2w).
struct W2C { double m[16]; };

__global__ void pgn_cyl_bbox_kernel(const float* __restrict__ cyls, int n, W2C w2c, int H, int W, double focal, int* __restrict__ bbox) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const double cx = cyls[p * 5], cz = cyls[p * 5 + 1], R = cyls[p * 5 + 2];
  const double ys[2] = {(double)cyls[p * 5 + 3], (double)cyls[p * 5 + 4]};
  double umin = 1e300, umax = -1e300, vmin = 1e300, vmax = -1e300;
  const double step = (2.0 * 3.141592653589793) / 49.0;          // np.linspace(0., 2 * np.pi, 50)
  for (int i = 0; i < 50; ++i) {
    const double ang = i == 49 ? 2.0 * 3.141592653589793 : i * step;
    const double x = cx + cos(ang) * R, z = cz + sin(ang) * R;
    for (int c = 0; c < 2; ++c) {
      const double y = ys[c];
      const double X = x * w2c.m[0] + y * w2c.m[1] + z * w2c.m[2] + w2c.m[3];
      const double Y = x * w2c.m[4] + y * w2c.m[5] + z * w2c.m[6] + w2c.m[7];
      const double Z = x * w2c.m[8] + y * w2c.m[9] + z * w2c.m[10] + w2c.m[11];
      const double u = (X * focal) / Z, v = (Y * focal) / Z;
      umin = fmin(umin, u); umax = fmax(umax, u); vmin = fmin(vmin, v); vmax = fmax(vmax, v);
    }
  }
  const int ox = (int)(W * 0.5), oy = (int)(H * 0.5);
  int x0 = (int)floor(umin) + ox, y0 = (int)floor(vmin) + oy, x1 = (int)ceil(umax) + ox, y1 = (int)ceil(vmax) + oy;
  x0 = min(max(x0, 0), W - 1); x1 = min(max(x1, 0), W - 1);
  y0 = min(max(y0, 0), H - 1); y1 = min(max(y1, 0), H - 1);
  bbox[p * 4] = x0; bbox[p * 4 + 1] = y0; bbox[p * 4 + 2] = x1; bbox[p * 4 + 3] = y1;
}

// get_rays restricted to each pose's bbox (rows [y0,y1) x cols [x0,x1)), written back to back:
// ray (pose p, local i) lands at row offsets[p] + i.  blockIdx.y = pose, grid-stride over its rays.
__global__ void pgn_generate_rays_batch_kernel(int H, int W, float focal, const float* __restrict__ c2w, const int* __restrict__ bbox,
                                               const long long* __restrict__ offsets, float* __restrict__ rb, int* __restrict__ pose_idx) {
  const int p = blockIdx.y;
  const int x0 = bbox[p * 4], y0 = bbox[p * 4 + 1], bw = bbox[p * 4 + 2] - x0;
  const long long base = offsets[p], n = offsets[p + 1] - base;
  const float c0 = c2w[0], c1 = c2w[1], c2 = c2w[2], c4 = c2w[4], c5 = c2w[5], c6 = c2w[6], c8 = c2w[8], c9 = c2w[9], c10 = c2w[10];
  const float ox = c2w[3], oy = c2w[7], oz = c2w[11];
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    const int px = x0 + (int)(idx % bw), py = y0 + (int)(idx / bw);
    const float dx = __fdiv_rn(__fsub_rn((float)px, W * 0.5f), focal);
    const float dy = -__fdiv_rn(__fsub_rn((float)py, H * 0.5f), focal);
    const float dz = -1.0f;
    // torch.sum(dirs[..., None, :] * c2w[:3,:3], -1): ((dx*c0 + dy*c1) + dz*c2), no FMA contraction
    const float d0 = __fadd_rn(__fadd_rn(__fmul_rn(dx, c0), __fmul_rn(dy, c1)), __fmul_rn(dz, c2));
    const float d1 = __fadd_rn(__fadd_rn(__fmul_rn(dx, c4), __fmul_rn(dy, c5)), __fmul_rn(dz, c6));
    const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, c8), __fmul_rn(dy, c9)), __fmul_rn(dz, c10));
    const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2)));
    float* o = rb + (base + idx) * 11;
    o[0] = ox; o[1] = oy; o[2] = oz; o[3] = d0; o[4] = d1; o[5] = d2; o[6] = 0.f; o[7] = 1.f;
    o[8] = __fdiv_rn(d0, nrm); o[9] = __fdiv_rn(d1, nrm); o[10] = __fdiv_rn(d2, nrm);
    pose_idx[base + idx] = p;
  }
}

// run_nerf.py:100-133 for B frames at once: image[p] = bg; image[p][bbox] = rgb + (1 - acc) * bg
__global__ void pgn_compose_frames_batch_kernel(int H, int W, const int* __restrict__ bbox, const long long* __restrict__ offsets,
                                                const float* __restrict__ rgb, const float* __restrict__ acc, float bg,
                                                float* __restrict__ images) {
  const int p = blockIdx.y;
  const int x0 = bbox[p * 4], y0 = bbox[p * 4 + 1], x1 = bbox[p * 4 + 2], y1 = bbox[p * 4 + 3], bw = x1 - x0;
  const long long base = offsets[p];
  const long long n = (long long)H * W;
  float* img = images + (size_t)p * n * 3;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
    const int px = (int)(q % W), py = (int)(q / W);
    float r = bg, g = bg, b = bg;
    if (px >= x0 && px < x1 && py >= y0 && py < y1) {
      const long long i = base + (long long)(py - y0) * bw + (px - x0);
      const float back = __fmul_rn(__fsub_rn(1.0f, acc[i]), bg);
      r = __fadd_rn(rgb[i * 3], back); g = __fadd_rn(rgb[i * 3 + 1], back); b = __fadd_rn(rgb[i * 3 + 2], back);
    }
    img[q * 3] = r; img[q * 3 + 1] = g; img[q * 3 + 2] = b;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Backward of pgn_pose_fk_kernel (pgn_stage_kernels.cu): one thread per pose, fp64.
//   forward:  Rg_i = Rg_p R_i,  T_i = Rg_p off_i + T_p,  skts_i = [Rg_i^T | -Rg_i^T T_i]
//   backward: dRg_i = GS_i^T - T_i gs_i^T,  dT_i = -Rg_i gs_i;  leaves -> root:
//             dRg_p += dRg_i R_i^T + dT_i off_i^T,  dT_p += dT_i,  dR_i = Rg_p^T dRg_i;
//             dL/d bones_i = <dR_i, dR/dr> with dR/dr_j = (r_j [r]x + [r x (I - R) e_j]x) R / |r|^2
//             (Gallego & Yezzi 2015; [e_j]x at r = 0).
// g_kps (optional) adds dL/d kps = dL/dT directly (the key points are FK outputs too).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void rodrigues(const double r[3], double R[9]) {
  const double th = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  if (th < 1e-12) { R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1; return; }
  const double kx = r[0] / th, ky = r[1] / th, kz = r[2] / th, s = sin(th), c = cos(th), v = 1.0 - c;
  R[0] = c + kx * kx * v;      R[1] = kx * ky * v - kz * s; R[2] = kx * kz * v + ky * s;
  R[3] = ky * kx * v + kz * s; R[4] = c + ky * ky * v;      R[5] = ky * kz * v - kx * s;
  R[6] = kz * kx * v - ky * s; R[7] = kz * ky * v + kx * s; R[8] = c + kz * kz * v;
}

__global__ void pgn_pose_fk_backward_kernel(const float* __restrict__ bones, const float* __restrict__ rest, int n_poses,
                                            const float* __restrict__ g_skts, const float* __restrict__ g_kps,
                                            float* __restrict__ g_bones) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_poses) return;
  double Rl[PGN_J][9], Rg[PGN_J][9], T[PGN_J][3], dRg[PGN_J][9], dT[PGN_J][3];
  for (int i = 0; i < PGN_J; ++i) {
    const double r[3] = {(double)bones[(p * PGN_J + i) * 3], (double)bones[(p * PGN_J + i) * 3 + 1], (double)bones[(p * PGN_J + i) * 3 + 2]};
    rodrigues(r, Rl[i]);
    if (i == 0) {
      for (int k = 0; k < 9; ++k) Rg[0][k] = Rl[0][k];
      for (int k = 0; k < 3; ++k) T[0][k] = rest[k];
    } else {
      const int pa = c_parents[i];
      const double d[3] = {(double)rest[i * 3] - (double)rest[pa * 3], (double)rest[i * 3 + 1] - (double)rest[pa * 3 + 1],
                           (double)rest[i * 3 + 2] - (double)rest[pa * 3 + 2]};
      for (int a = 0; a < 3; ++a) {
        for (int b = 0; b < 3; ++b)
          Rg[i][a * 3 + b] = Rg[pa][a * 3] * Rl[i][b] + Rg[pa][a * 3 + 1] * Rl[i][3 + b] + Rg[pa][a * 3 + 2] * Rl[i][6 + b];
        T[i][a] = Rg[pa][a * 3] * d[0] + Rg[pa][a * 3 + 1] * d[1] + Rg[pa][a * 3 + 2] * d[2] + T[pa][a];
      }
    }
  }
  for (int i = 0; i < PGN_J; ++i) {
    const float* g = g_skts + ((size_t)p * PGN_J + i) * 16;
    const double gs[3] = {(double)g[3], (double)g[7], (double)g[11]};
    for (int a = 0; a < 3; ++a) {
      for (int b = 0; b < 3; ++b) dRg[i][a * 3 + b] = (double)g[b * 4 + a] - T[i][a] * gs[b];     // GS^T - T gs^T
      dT[i][a] = -(Rg[i][a * 3] * gs[0] + Rg[i][a * 3 + 1] * gs[1] + Rg[i][a * 3 + 2] * gs[2]);
      if (g_kps) dT[i][a] += (double)g_kps[(p * PGN_J + i) * 3 + a];
    }
  }
  for (int i = PGN_J - 1; i >= 0; --i) {
    double dRl[9];
    if (i == 0) {
      for (int k = 0; k < 9; ++k) dRl[k] = dRg[0][k];
    } else {
      const int pa = c_parents[i];
      const double d[3] = {(double)rest[i * 3] - (double)rest[pa * 3], (double)rest[i * 3 + 1] - (double)rest[pa * 3 + 1],
                           (double)rest[i * 3 + 2] - (double)rest[pa * 3 + 2]};
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
          // dRg_p += dRg_i R_i^T + dT_i off^T
          dRg[pa][a * 3 + b] += dRg[i][a * 3] * Rl[i][b * 3] + dRg[i][a * 3 + 1] * Rl[i][b * 3 + 1] + dRg[i][a * 3 + 2] * Rl[i][b * 3 + 2] +
                                dT[i][a] * d[b];
          // dR_i = Rg_p^T dRg_i
          dRl[a * 3 + b] = Rg[pa][a] * dRg[i][b] + Rg[pa][3 + a] * dRg[i][3 + b] + Rg[pa][6 + a] * dRg[i][6 + b];
        }
      for (int a = 0; a < 3; ++a) dT[pa][a] += dT[i][a];
    }
    const double r[3] = {(double)bones[(p * PGN_J + i) * 3], (double)bones[(p * PGN_J + i) * 3 + 1], (double)bones[(p * PGN_J + i) * 3 + 2]};
    const double th2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
    const double* R = Rl[i];
    for (int j = 0; j < 3; ++j) {
      double D[9];                                       // dR / dr_j
      if (th2 < 1e-24) {
        for (int k = 0; k < 9; ++k) D[k] = 0.0;
        if (j == 0) { D[5] = -1; D[7] = 1; } else if (j == 1) { D[2] = 1; D[6] = -1; } else { D[1] = -1; D[3] = 1; }
      } else {
        // u = r x ((I - R) e_j)
        const double c[3] = {(j == 0 ? 1.0 : 0.0) - R[j], (j == 1 ? 1.0 : 0.0) - R[3 + j], (j == 2 ? 1.0 : 0.0) - R[6 + j]};
        const double u[3] = {r[1] * c[2] - r[2] * c[1], r[2] * c[0] - r[0] * c[2], r[0] * c[1] - r[1] * c[0]};
        // M = r_j [r]x + [u]x
        const double M[9] = {0.0, -(r[j] * r[2] + u[2]), r[j] * r[1] + u[1],
                             r[j] * r[2] + u[2], 0.0, -(r[j] * r[0] + u[0]),
                             -(r[j] * r[1] + u[1]), r[j] * r[0] + u[0], 0.0};
        for (int a = 0; a < 3; ++a)
          for (int b = 0; b < 3; ++b)
            D[a * 3 + b] = (M[a * 3] * R[b] + M[a * 3 + 1] * R[3 + b] + M[a * 3 + 2] * R[6 + b]) / th2;
      }
      double acc = 0.0;
      for (int k = 0; k < 9; ++k) acc += dRl[k] * D[k];
      g_bones[(p * PGN_J + i) * 3 + j] = (float)acc;
    }
  }
}

}  // namespace

cudaError_t pgn_launch_cyl_bboxes(const float* cyls, int n, const double* w2c16, int H, int W, double focal, int* bbox, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  W2C m;
  for (int i = 0; i < 16; ++i) m.m[i] = w2c16[i];
  pgn_cyl_bbox_kernel<<<(n + 63) / 64, 64, 0, stream>>>(cyls, n, m, H, W, focal, bbox);
  return cudaGetLastError();
}

cudaError_t pgn_launch_generate_rays_batch(int H, int W, float focal, const float* c2w12_dev, const int* bbox, const long long* offsets,
                                           int n_poses, long long max_rays_per_pose, float* ray_batch, int* pose_idx, cudaStream_t stream) {
  if (n_poses <= 0 || max_rays_per_pose <= 0) return cudaSuccess;
  const int block = 256;
  const long long per = (max_rays_per_pose + block - 1) / block;
  dim3 grid((unsigned)(per < 148 ? per : 148), (unsigned)n_poses);
  pgn_generate_rays_batch_kernel<<<grid, block, 0, stream>>>(H, W, focal, c2w12_dev, bbox, offsets, ray_batch, pose_idx);
  return cudaGetLastError();
}

cudaError_t pgn_launch_compose_frames_batch(int H, int W, const int* bbox, const long long* offsets, int n_poses, const float* rgb,
                                            const float* acc, float bg, float* images, cudaStream_t stream) {
  if (n_poses <= 0) return cudaSuccess;
  const int block = 256;
  const long long per = ((long long)H * W + block - 1) / block;
  dim3 grid((unsigned)(per < 148 ? per : 148), (unsigned)n_poses);
  pgn_compose_frames_batch_kernel<<<grid, block, 0, stream>>>(H, W, bbox, offsets, rgb, acc, bg, images);
  return cudaGetLastError();
}

// Rows of selected rays out of per-sample planes (the GAN step's backward walks the rays the HMR crop reads in chunks):
// dst[p][i] = src[p][idx[i]] for n_planes planes of rows of row_bytes bytes (a multiple of 16; 16-byte aligned planes).
// What `index_select` does at 150-230 us per call on these shapes (run_gan.py has no counterpart: its render is not
// differentiated); here one launch per tensor at HBM speed.
// An index outside the source plane writes zeros and latches status code 950 (pgn_check_device_status reports it).
__global__ void pgn_gather_ray_rows_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, const long long* __restrict__ idx,
                                           long long n_idx, int row_vec, int n_planes, long long src_plane_vec, long long dst_plane_vec,
                                           int* __restrict__ status) {
  const long long total = (long long)n_planes * n_idx * row_vec;
  const long long src_rows = src_plane_vec / row_vec;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long rowi = i / row_vec;
    const int v = (int)(i - rowi * row_vec);
    const long long pl = rowi / n_idx, r = rowi - pl * n_idx;
    const long long sr = __ldg(idx + r);
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (sr >= 0 && sr < src_rows) val = __ldg(src + pl * src_plane_vec + sr * row_vec + v);
    else if (v == 0 && pl == 0) *status = 950;
    dst[pl * dst_plane_vec + r * row_vec + v] = val;
  }
}

cudaError_t pgn_launch_gather_ray_rows(const void* src, void* dst, const long long* idx, long long n_idx, long long row_bytes,
                                       int n_planes, long long src_plane_bytes, long long dst_plane_bytes, int* status, int num_sms,
                                       cudaStream_t stream) {
  if (n_idx <= 0 || n_planes <= 0) return cudaSuccess;
  const long long total = (long long)n_planes * n_idx * (row_bytes / 16);
  const long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms * 8;
  pgn_gather_ray_rows_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, stream>>>(
      reinterpret_cast<const uint4*>(src), reinterpret_cast<uint4*>(dst), idx, n_idx, (int)(row_bytes / 16), n_planes,
      src_plane_bytes / 16, dst_plane_bytes / 16, status);
  return cudaGetLastError();
}

cudaError_t pgn_launch_pose_fk_backward(const float* bones, const float* rest, int n_poses, const float* g_skts, const float* g_kps,
                                        float* g_bones, cudaStream_t stream) {
  if (n_poses <= 0) return cudaSuccess;
  pgn_pose_fk_backward_kernel<<<(n_poses + 31) / 32, 32, 0, stream>>>(bones, rest, n_poses, g_skts, g_kps, g_bones);
  return cudaGetLastError();
}
