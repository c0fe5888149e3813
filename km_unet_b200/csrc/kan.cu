// kan.cu -- C-ABI entry points of the KANConv2d / KANLinear op (argument checks + kernel-family dispatch).
#include "common.cuh"
#include "kan_common.cuh"

using namespace kmu;
using namespace kmu::kan;

static int check_desc(const kmu_kanconv2d_desc* d, const char* who) {
  KMU_REQUIRE(d != nullptr, KMU_ERR_BAD_ARG, "%s: null descriptor", who);
  KMU_REQUIRE(d->B > 0 && d->Cin > 0 && d->H > 0 && d->W > 0 && d->Cout > 0, KMU_ERR_BAD_ARG,
              "%s: non-positive shape B=%d Cin=%d H=%d W=%d Cout=%d", who, d->B, d->Cin, d->H, d->W, d->Cout);
  KMU_REQUIRE(d->ksize > 0 && d->stride > 0 && d->padding >= 0, KMU_ERR_BAD_ARG, "%s: bad ksize/stride/padding %d/%d/%d", who,
              d->ksize, d->stride, d->padding);
  KMU_REQUIRE(d->H + 2 * d->padding >= d->ksize && d->W + 2 * d->padding >= d->ksize, KMU_ERR_BAD_ARG,
              "%s: kernel %d larger than padded input %dx%d", who, d->ksize, d->H + 2 * d->padding, d->W + 2 * d->padding);
  KMU_REQUIRE(d->spline_order == ORDER && d->grid_size + d->spline_order == NB, KMU_ERR_UNSUPPORTED,
              "%s: only cubic splines with 8 basis functions (grid_size=5, spline_order=3) are implemented, got %d/%d", who,
              d->grid_size, d->spline_order);
  KMU_REQUIRE(d->precision == KMU_PREC_FP32 || d->precision == KMU_PREC_BF16, KMU_ERR_BAD_ARG, "%s: bad precision %d", who,
              d->precision);
  return KMU_OK;
}

extern "C" {

int kmu_kanconv2d_path(const kmu_kanconv2d_desc* d) {
  if (!d) return 0;
  return (d->precision == KMU_PREC_BF16 && tc::supported(*d)) ? 1 : 0;
}

size_t kmu_kanconv2d_fwd_workspace_bytes(const kmu_kanconv2d_desc* d) {
  if (check_desc(d, "kanconv2d_fwd_workspace_bytes") != KMU_OK) return 0;
  if (kmu_kanconv2d_path(d)) return tc::fwd_workspace(make_dims(*d));
  return simt_fwd_workspace(make_dims(*d));
}

size_t kmu_kanconv2d_bwd_workspace_bytes(const kmu_kanconv2d_desc* d) {
  if (check_desc(d, "kanconv2d_bwd_workspace_bytes") != KMU_OK) return 0;
  if (kmu_kanconv2d_path(d)) return tc::bwd_workspace(make_dims(*d));
  return simt_bwd_workspace(make_dims(*d));
}

int kmu_kanconv2d_fwd(const kmu_kanconv2d_fwd_args* a, kmu_stream stream) {
  KMU_REQUIRE(a != nullptr, KMU_ERR_BAD_ARG, "kanconv2d_fwd: null args");
  int st = check_desc(&a->d, "kanconv2d_fwd");
  if (st != KMU_OK) return st;
  KMU_REQUIRE(a->x && a->base_weight && a->spline_weight && a->grid && a->y, KMU_ERR_BAD_ARG, "kanconv2d_fwd: null tensor");
  KMU_REQUIRE(!a->d.has_scaler || a->spline_scaler, KMU_ERR_BAD_ARG, "kanconv2d_fwd: has_scaler set but spline_scaler is null");
  kmu_kanconv2d_fwd_args b = *a;
  if (!b.d.has_scaler) b.spline_scaler = nullptr;
  if (kmu_kanconv2d_path(&a->d)) {
    KMU_REQUIRE(kmu_device_supported(), KMU_ERR_DEVICE, "kanconv2d_fwd: the tcgen05 family needs an sm_100 device");
    return tc::forward(&b, make_dims(a->d), (cudaStream_t)stream);
  }
  return simt_forward(&b, make_dims(a->d), (cudaStream_t)stream);
}

int kmu_kanconv2d_bwd(const kmu_kanconv2d_bwd_args* a, kmu_stream stream) {
  KMU_REQUIRE(a != nullptr, KMU_ERR_BAD_ARG, "kanconv2d_bwd: null args");
  int st = check_desc(&a->d, "kanconv2d_bwd");
  if (st != KMU_OK) return st;
  KMU_REQUIRE(a->x && a->dy && a->base_weight && a->spline_weight && a->grid, KMU_ERR_BAD_ARG, "kanconv2d_bwd: null tensor");
  KMU_REQUIRE(!a->d.has_scaler || a->spline_scaler, KMU_ERR_BAD_ARG, "kanconv2d_bwd: has_scaler set but spline_scaler is null");
  KMU_REQUIRE(!a->d_base_weight || a->d_spline_weight, KMU_ERR_BAD_ARG, "kanconv2d_bwd: d_spline_weight is null");
  KMU_REQUIRE(!a->d_base_weight || !a->d.has_scaler || a->d_spline_scaler, KMU_ERR_BAD_ARG,
              "kanconv2d_bwd: d_spline_scaler is null");
  kmu_kanconv2d_bwd_args b = *a;
  if (!b.d.has_scaler) { b.spline_scaler = nullptr; b.d_spline_scaler = nullptr; }
  if (kmu_kanconv2d_path(&a->d)) {
    KMU_REQUIRE(kmu_device_supported(), KMU_ERR_DEVICE, "kanconv2d_bwd: the tcgen05 family needs an sm_100 device");
    return tc::backward(&b, make_dims(a->d), (cudaStream_t)stream);
  }
  return simt_backward(&b, make_dims(a->d), (cudaStream_t)stream);
}

void kmu_debug_flags(int flags) { tc::set_debug_flags(flags); }

}  // extern "C"
