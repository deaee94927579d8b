// Library-level entry points.
#include <stdlib.h>
#include "common.cuh"
extern "C" MDM_API int mdm_num_sms(void) {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n;
}
// Programmatic dependent launch switch (common.cuh): on by default; MDM_B200_PDL=0 or mdm_set_pdl(0) turns the launch
// attribute off (the kernels' griddepcontrol instructions are then no-ops).
static int g_pdl = -1;
int mdm_pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("MDM_B200_PDL");
    g_pdl = (e && e[0] == '0') ? 0 : 1;
  }
  return g_pdl;
}
extern "C" MDM_API int mdm_set_pdl(int on) {
  const int was = mdm_pdl_enabled();
  if (on >= 0) g_pdl = on ? 1 : 0;
  return was;
}
extern "C" MDM_API const char* mdm_version(void) { return "mdm_b200 0.2 (sm_100a)"; }
extern "C" MDM_API int mdm_sizeof_gemm_epi(void) { return (int)sizeof(MdmGemmEpi); }
extern "C" MDM_API int mdm_sizeof_rowop(void) { return (int)sizeof(MdmRowOp); }
extern "C" MDM_API int mdm_sizeof_ep_peers(void) { return (int)sizeof(MdmEpPeers); }
extern "C" MDM_API int mdm_sizeof_bgemm(void) { return (int)sizeof(MdmBgemm); }
