// GPU input pipeline: the per-volume pre-processing of scripts/data.py:92-162 (and its twin
// scripts/data_inference_nii.py:96-176) as ONE kernel, from the raw NIfTI voxel array to the (1, D, H, W) tensor the
// patch-embed kernel consumes:
//   hu = slope * raw + intercept                                   (data.py:109)
//   (X, Y, Z) -> (Z, X, Y); F.interpolate(trilinear, align_corners=False) to
//        (int(Z * z_spacing / 1.5), int(X * xy_spacing / 0.75), int(Y * xy_spacing / 0.75))      (data.py:12-34, :111-117)
//   back to (X', Y', Z'); clip to [-1000, 1000]; / 1000                                            (data.py:118-123)
//   centre crop / pad (value -1) to (480, 480, 240); permute to (Z, X, Y) = (240, 480, 480)        (data.py:127-160)
// The reference does this on the CPU in float64 (nibabel get_fdata) at ~1 s per volume; here every OUTPUT voxel gathers its
// 8 source voxels (the interpolation weights sum to one, so applying slope/intercept after the interpolation is the same
// affine map). Output: fp32 in [-1, 1] (reference contract) or int16 HU (rounded; what the patch-embed kernel reads as x/1000
// -- half the bytes of the reference's fp32 and 0.5 HU = 5e-4 of quantisation).
// HBM-bound: reads <= raw bytes (each raw voxel is touched by ~ (1/scale)^3 * 8 output gathers, L2-resident), writes the output once.
#include "common.cuh"
#include "../../include/ctclip_b200.h"

namespace ctb {

struct PreArgs {
  const void* raw;
  int raw_dtype;            // 0 = float32, 1 = int16
  int X, Y, Z;              // raw array dims, C-contiguous (x, y, z) as nibabel's get_fdata() returns it
  float slope, intercept;
  int Zr, Xr, Yr;           // resized dims
  float sz, sx, sy;         // in/out ratios of the interpolation (PyTorch: scale = in_size / out_size)
  int D, H, W;              // output dims (frames, height, width) = (240, 480, 480)
  int d_pb, h_pb, w_pb;        // pad-before of each output axis (data.py:149-156)
  int d_len, h_len, w_len;     // cropped extents
  int d_st, h_st, w_st;        // crop start in the resized volume (data.py:137-142)
  void* out;
  int out_dtype;            // 0 = float32 in [-1, 1], 1 = int16 HU
  float pad_value;          // -1 (data.py:158)
};

__device__ __forceinline__ void src_index(float scale, int dst, int in_size, int& i0, int& i1, float& l1) {
  // aten/src/ATen/native/UpSample.h area_pixel_compute_source_index (align_corners = False, cubic = False)
  float src = scale * ((float)dst + 0.5f) - 0.5f;
  if (src < 0.f) src = 0.f;
  i0 = (int)src;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
  l1 = src - (float)i0;
}

template <typename T>
__global__ void __launch_bounds__(256) ct_preprocess_kernel(PreArgs a) {
  const T* __restrict__ raw = reinterpret_cast<const T*>(a.raw);
  const long long total = (long long)a.D * a.H * a.W;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(idx % a.W);
    const int h = (int)((idx / a.W) % a.H);
    const int d = (int)(idx / ((long long)a.W * a.H));
    const int dz = d - a.d_pb, dx = h - a.h_pb, dy = w - a.w_pb;          // position inside the cropped block
    float val = a.pad_value;
    if (dz >= 0 && dz < a.d_len && dx >= 0 && dx < a.h_len && dy >= 0 && dy < a.w_len) {
      const int zr = dz + a.d_st, xr = dx + a.h_st, yr = dy + a.w_st;     // indices in the resized (Z', X', Y') volume
      int z0, z1, x0, x1, y0, y1;
      float lz, lx, ly;
      src_index(a.sz, zr, a.Z, z0, z1, lz);
      src_index(a.sx, xr, a.X, x0, x1, lx);
      src_index(a.sy, yr, a.Y, y0, y1, ly);
      auto at = [&](int x, int y, int z) { return (float)raw[((long long)x * a.Y + y) * a.Z + z]; };
      const float c00 = at(x0, y0, z0) * (1.f - ly) + at(x0, y1, z0) * ly;
      const float c01 = at(x1, y0, z0) * (1.f - ly) + at(x1, y1, z0) * ly;
      const float c10 = at(x0, y0, z1) * (1.f - ly) + at(x0, y1, z1) * ly;
      const float c11 = at(x1, y0, z1) * (1.f - ly) + at(x1, y1, z1) * ly;
      const float c0 = c00 * (1.f - lx) + c01 * lx, c1 = c10 * (1.f - lx) + c11 * lx;
      float hu = (c0 * (1.f - lz) + c1 * lz) * a.slope + a.intercept;
      hu = fminf(fmaxf(hu, -1000.f), 1000.f);
      val = hu * (1.0f / 1000.0f);
    }
    if (a.out_dtype == 0) reinterpret_cast<float*>(a.out)[idx] = val;
    else reinterpret_cast<int16_t*>(a.out)[idx] = (int16_t)__float2int_rn(val * 1000.f);
  }
}

}  // namespace ctb

using namespace ctb;

extern "C" int ctclip_ct_preprocess(const ctclip_preprocess_args* p, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTB_CHECK_ARG(p && p->raw && p->out, "ct_preprocess: null pointer");
  CTB_CHECK_ARG(p->raw_dtype == 0 || p->raw_dtype == 1, "ct_preprocess: raw_dtype must be 0 (float32) or 1 (int16)");
  CTB_CHECK_ARG(p->out_dtype == 0 || p->out_dtype == 1, "ct_preprocess: out_dtype must be 0 (float32) or 1 (int16 HU)");
  CTB_CHECK_ARG(p->X > 0 && p->Y > 0 && p->Z > 0 && p->out_d > 0 && p->out_h > 0 && p->out_w > 0, "ct_preprocess: bad dims");
  CTB_CHECK_ARG(p->xy_spacing > 0. && p->z_spacing > 0. && p->target_xy > 0. && p->target_z > 0., "ct_preprocess: bad spacing");
  PreArgs a;
  a.raw = p->raw; a.raw_dtype = p->raw_dtype; a.X = p->X; a.Y = p->Y; a.Z = p->Z;
  a.slope = p->slope; a.intercept = p->intercept;
  // data.py:26-31: new_shape = int(original * current / target), evaluated in double like Python floats
  a.Zr = (int)((double)p->Z * (p->z_spacing / p->target_z));
  a.Xr = (int)((double)p->X * (p->xy_spacing / p->target_xy));
  a.Yr = (int)((double)p->Y * (p->xy_spacing / p->target_xy));
  CTB_CHECK_ARG(a.Zr > 0 && a.Xr > 0 && a.Yr > 0, "ct_preprocess: resized volume is empty");
  a.sz = (float)((double)p->Z / (double)a.Zr); a.sx = (float)((double)p->X / (double)a.Xr); a.sy = (float)((double)p->Y / (double)a.Yr);
  a.D = p->out_d; a.H = p->out_h; a.W = p->out_w;
  auto crop = [](int have, int want, int& pad_before, int& len, int& start) {   // data.py:133-156
    start = (have - want) / 2 > 0 ? (have - want) / 2 : 0;
    const int end = start + want < have ? start + want : have;
    len = end - start;
    pad_before = (want - len) / 2;
  };
  crop(a.Xr, a.H, a.h_pb, a.h_len, a.h_st);
  crop(a.Yr, a.W, a.w_pb, a.w_len, a.w_st);
  crop(a.Zr, a.D, a.d_pb, a.d_len, a.d_st);
  a.out = p->out; a.out_dtype = p->out_dtype; a.pad_value = p->pad_value;
  const long long total = (long long)a.D * a.H * a.W;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (a.raw_dtype == 0) ct_preprocess_kernel<float><<<(int)blocks, 256, 0, stream>>>(a);
  else ct_preprocess_kernel<int16_t><<<(int)blocks, 256, 0, stream>>>(a);
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}
