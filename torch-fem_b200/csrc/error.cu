// Error reporting of the C ABI (mirrors AMGX_get_error_string, /root/reference/src/torchfem/amgx.py:195-201).
#include <string.h>

#include "common.cuh"

namespace tfem {
static thread_local char g_last[512] = "";

void set_last_error(const char* what, const char* detail) {
  snprintf(g_last, sizeof(g_last), "%s: %s", what ? what : "", detail ? detail : "");
}
}  // namespace tfem

extern "C" int tfem_version(void) { return 100; /* 0.1.0 */ }

extern "C" int tfem_get_error_string(int rc, char* buf, int len) {
  if (!buf || len <= 0) return TFEM_ERR_INVALID;
  const char* base = "unknown error";
  switch (rc) {
    case TFEM_OK: base = "success"; break;
    case TFEM_ERR_INVALID: base = "invalid argument"; break;
    case TFEM_ERR_CUDA: base = "CUDA runtime error"; break;
    case TFEM_ERR_CAPACITY: base = "capacity limit exceeded"; break;
    case TFEM_ERR_NOT_CONVERGED: base = "Krylov solver did not converge"; break;
    case TFEM_ERR_BREAKDOWN: base = "Krylov solver breakdown"; break;
    case TFEM_ERR_NCCL: base = "NCCL error"; break;
    case TFEM_ERR_COMM: base = "peer communication failed"; break;
  }
  if (rc != TFEM_OK && tfem::g_last[0])
    snprintf(buf, (size_t)len, "%s (%s)", base, tfem::g_last);
  else
    snprintf(buf, (size_t)len, "%s", base);
  return TFEM_OK;
}
