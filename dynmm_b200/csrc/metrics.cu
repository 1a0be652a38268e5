// Evaluation post-processing on the device (FusionDynMM/eval.py:117-141 and
// src/confusion_matrix.py:85-178): arg-max over classes, void masking (label 0), confusion
// matrix by bincount of num_classes*label + pred, IoU / mIoU in double precision.
// Integer work: bit-exact with the reference.
#include "common.cuh"

namespace dynmm {
namespace {

__global__ void argmax_confusion_kernel(const float* __restrict__ logits, const uint8_t* __restrict__ label_orig,
                                        int c, long long hw, long long total, long long* __restrict__ cm,
                                        uint8_t* __restrict__ pred_out) {
  extern __shared__ int s_cm[];      // [c*c]
  for (int i = threadIdx.x; i < c * c; i += blockDim.x) s_cm[i] = 0;
  __syncthreads();
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    const long long n = i / hw, p = i - n * hw;
    const float* x = logits + n * c * hw + p;
    float best = x[0];
    int arg = 0;
    for (int k = 1; k < c; ++k) {
      const float v = __ldg(x + k * hw);
      if (v > best) {          // strict: first maximum wins, like torch.argmax
        best = v;
        arg = k;
      }
    }
    if (pred_out) pred_out[i] = static_cast<uint8_t>(arg);
    if (label_orig) {
      const int lab = label_orig[i];
      if (lab > 0 && lab <= c) atomicAdd(&s_cm[(lab - 1) * c + arg], 1);   // void (0) is ignored; label -= 1
    }
  }
  __syncthreads();
  if (cm) {
    for (int i = threadIdx.x; i < c * c; i += blockDim.x) {
      const int v = s_cm[i];
      if (v) atomicAdd(reinterpret_cast<unsigned long long*>(cm + i), static_cast<unsigned long long>(v));
    }
  }
}

__global__ void miou_kernel(const long long* __restrict__ cm, int c, double* __restrict__ iou, double* __restrict__ miou) {
  extern __shared__ double s_iou[];
  for (int k = threadIdx.x; k < c; k += blockDim.x) {
    double row = 0, col = 0;
    for (int j = 0; j < c; ++j) {
      row += (double)cm[k * c + j];
      col += (double)cm[j * c + k];
    }
    const double d = (double)cm[k * c + k];
    const double v = d / (row + col - d + 1e-15);      // confusion_matrix.py:153
    s_iou[k] = v;
    if (iou) iou[k] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    for (int k = 0; k < c; ++k) s += s_iou[k];
    *miou = s / c;
  }
}

}  // namespace
}  // namespace dynmm

using namespace dynmm;

extern "C" int dynmm_argmax_confusion(const float* logits, const uint8_t* label_orig, int n, int c, int h, int w,
                                      long long* cm, uint8_t* pred_out, void* stream) {
  DYNMM_CHECK_ARG(logits && n >= 1 && c >= 1 && c <= 96 && h >= 1 && w >= 1, "argmax_confusion: bad args (c <= 96)");
  DYNMM_CHECK_ARG((label_orig != nullptr) == (cm != nullptr), "argmax_confusion: labels and matrix come together");
  const long long hw = 1LL * h * w, total = hw * n;
  long long blocks = (total + 255) / 256;
  const long long cap = 4LL * num_sms();
  if (blocks > cap) blocks = cap;
  argmax_confusion_kernel<<<(int)blocks, 256, c * c * sizeof(int), static_cast<cudaStream_t>(stream)>>>(
      logits, label_orig, c, hw, total, cm, pred_out);
  DYNMM_LAUNCH_CHECK();
  return DYNMM_OK;
}

extern "C" int dynmm_miou(const long long* cm, int c, double* iou, double* miou, void* stream) {
  DYNMM_CHECK_ARG(cm && miou && c >= 1 && c <= 1024, "miou: bad args");
  miou_kernel<<<1, 128, c * sizeof(double), static_cast<cudaStream_t>(stream)>>>(cm, c, iou, miou);
  DYNMM_LAUNCH_CHECK();
  return DYNMM_OK;
}
