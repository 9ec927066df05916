// multi_gpu.cu -- peer (NVLink) exchange for sharded particle systems.  Filled in by the multi-GPU milestone.
#include "engine.h"
using namespace mpl;
extern "C" int mpl_ps_peer_export(mpl_ps* ps, void* blob) { (void)ps; (void)blob; return fail(MPL_ERR_UNSUPPORTED, "peer exchange: not in this build"); }
extern "C" int mpl_ps_peer_attach(mpl_ps* ps, int rank, int world, const void* blobs) { (void)ps; (void)rank; (void)world; (void)blobs; return fail(MPL_ERR_UNSUPPORTED, "peer exchange: not in this build"); }
extern "C" int mpl_ps_peer_detach(mpl_ps* ps) { (void)ps; return fail(MPL_ERR_UNSUPPORTED, "peer exchange: not in this build"); }
