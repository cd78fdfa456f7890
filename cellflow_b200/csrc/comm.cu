// comm.cu — multi-GPU slab decomposition: halo + migration exchange over NCCL (dlopen'd).
// Round-1 state: entry points exist; the exchange itself lands with the slab engine.
#include <cstring>
#include "../../include/cellflow_b200.h"

extern "C" int cf_nccl_unique_id(void* id128) {
    if (!id128) return CF_ERR_ARG;
    memset(id128, 0, 128);
    return CF_ERR_NCCL;
}
extern "C" int cf_comm_init(cf_sim*, int, int, const void*, int) { return CF_ERR_NCCL; }
