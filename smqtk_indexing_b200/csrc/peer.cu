// Peer memory plumbing for the row-sharded descriptor table: CUDA IPC export / import of a rank's shard so
// that a kernel on one GPU can load candidate rows straight from the GPU that owns them (NVLink / NVSwitch),
// see sb_rerank_peer in rerank.cu.  One process per GPU: the exporter publishes the 64-byte handle of the
// allocation that holds its shard, every importer opens it WITH ITS OWN DEVICE CURRENT and
// cudaIpcMemLazyEnablePeerAccess, which is what makes the returned pointer loadable from the importer's
// kernels (a mapping opened under the exporter's device index is only reachable by copies, not by kernels).
#include <cuda.h>
#include <string.h>

#include "common.cuh"

extern "C" {

/* handle_out: 64 host bytes (cudaIpcMemHandle_t) of the ALLOCATION containing dev_ptr; offset_out = dev_ptr -
 * allocation base.  The allocation must come from cudaMalloc (torch's caching allocator without expandable
 * segments does) and stay alive while peers use it. */
int sb_ipc_export(const void* dev_ptr, void* handle_out, int64_t* offset_out) {
  SB_REQUIRE(dev_ptr && handle_out && offset_out, "sb_ipc_export: NULL pointer");
  typedef CUresult (*range_fn)(CUdeviceptr*, size_t*, CUdeviceptr);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  SB_CUDA_TRY(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qres));
  SB_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, "sb_ipc_export: cuMemGetAddressRange unavailable");
  CUdeviceptr base = 0;
  size_t size = 0;
  const CUresult r = reinterpret_cast<range_fn>(fn)(&base, &size, reinterpret_cast<CUdeviceptr>(dev_ptr));
  if (r != CUDA_SUCCESS) {
    sb::set_error("sb_ipc_export: cuMemGetAddressRange failed (%d)", (int)r);
    return SB_ERR_CUDA;
  }
  cudaIpcMemHandle_t h;
  SB_CUDA_TRY(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base)));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle_out, &h, sizeof(h));
  *offset_out = (int64_t)(reinterpret_cast<CUdeviceptr>(dev_ptr) - base);
  return SB_OK;
}

/* Opens a peer's handle under the CURRENT device; *base_out is the allocation base in this process (add the
 * exporter's offset).  Peer access between the two devices is enabled as part of the import. */
int sb_ipc_import(const void* handle, void** base_out) {
  SB_REQUIRE(handle && base_out, "sb_ipc_import: NULL pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  SB_CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *base_out = p;
  return SB_OK;
}

int sb_ipc_release(void* base) {
  if (base == nullptr) return SB_OK;
  SB_CUDA_TRY(cudaIpcCloseMemHandle(base));
  return SB_OK;
}

}  // extern "C"
