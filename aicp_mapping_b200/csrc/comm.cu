// comm.cu -- multi-GPU single registration (reading sharded across ranks, reference replicated).
// Filled in by the sharded-registration milestone; until then the entry points fail loudly.
#include "handle.cuh"

using namespace aicp;

extern "C" {

int aicp_b200_comm_unique_id(uint8_t id_out[128]) {
  (void)id_out;
  return fail(nullptr, AICP_B200_ERR_COMM, "sharded registration is not built into this library yet");
}

int aicp_b200_comm_init(aicp_b200_handle* hh, const uint8_t nccl_unique_id[128], int rank, int n_ranks) {
  (void)nccl_unique_id; (void)rank; (void)n_ranks;
  return fail(reinterpret_cast<Handle*>(hh), AICP_B200_ERR_COMM, "sharded registration is not built into this library yet");
}

int aicp_b200_comm_destroy(aicp_b200_handle* hh) {
  (void)hh;
  return AICP_B200_OK;
}

}
