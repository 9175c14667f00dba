// tcgen05 GEMM (placeholder until validated on hardware).
#include "car_common.cuh"
namespace car {
int launch_gemm_umma(const uint16_t *, const uint16_t *, int, const uint16_t *, const uint16_t *, int,
                     int, int, int, int, const GemmEpi &, const UmmaOut &, cudaStream_t) {
  set_error("tcgen05 GEMM not built");
  return -9;
}
}  // namespace car
extern "C" int car_gemm_umma_test(const uint16_t *, const uint16_t *, const uint16_t *, const uint16_t *,
                                  const float *, float *, int, int, int, int, int, void *) {
  car::set_error("tcgen05 GEMM not built");
  return -9;
}
