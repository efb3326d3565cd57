// temporary: launchers not yet implemented
#include "nxs_common.cuh"
namespace nxs {
int launch_istft(nxs_ctx*, const float2*, int64_t, int64_t, int64_t, const float*, int64_t, int64_t, int64_t, int,
                 double, float2*, cudaStream_t) { return NXS_EUNSUPPORTED; }
int launch_fir(nxs_ctx*, const float*, int64_t, int64_t, int64_t, const float*, int64_t, int, float*, int64_t,
               cudaStream_t) { return NXS_EUNSUPPORTED; }
}
