// Library-level entry points.
#include "common.cuh"
extern "C" MDM_API int mdm_num_sms(void) {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n;
}
extern "C" MDM_API const char* mdm_version(void) { return "mdm_b200 0.2 (sm_100a)"; }
extern "C" MDM_API int mdm_sizeof_gemm_epi(void) { return (int)sizeof(MdmGemmEpi); }
extern "C" MDM_API int mdm_sizeof_rowop(void) { return (int)sizeof(MdmRowOp); }
extern "C" MDM_API int mdm_sizeof_ep_peers(void) { return (int)sizeof(MdmEpPeers); }
extern "C" MDM_API int mdm_sizeof_bgemm(void) { return (int)sizeof(MdmBgemm); }
