// placeholder until K1 lands
#include "afs_internal.h"
namespace afs {
int tds_prepare_launch(const afs_tds_options *, int, int, const int32_t *, double, int, int, TdsLaunch *) { return AFS_ERR_UNSUPPORTED; }
cudaError_t launch_tds(const TdsLaunch &, const int32_t *, const afs_tube_frame *, const double *, double *, double *, uint32_t *, cudaStream_t) { return cudaErrorNotSupported; }
}
