// kernels_blo_generic.cuh - HOT LOOP B for any state count (amino acids): placeholder until the
// CTA-per-pair kernel lands; the DNA kernel lives in kernels_blo.cuh.
#pragma once
#include "kernels_blo.cuh"

namespace epa {
inline cudaError_t launch_blo_generic(int, int, int, size_t, int, const DevModel *, BloArgs &, void **, size_t *, cudaStream_t)
{
  return cudaErrorNotSupported;
}
}  // namespace epa
