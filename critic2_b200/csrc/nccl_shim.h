// nccl_shim.h -- NCCL is resolved lazily with dlopen so that the library has no link-time
// dependency on a particular libnccl: in a process that already loaded NCCL (e.g. the copy bundled
// with PyTorch, used by bench.py for torch.distributed plumbing) dlopen("libnccl.so.2") returns that
// same copy; otherwise the system library is used.  Single-GPU runs never touch NCCL.
#pragma once
#include <nccl.h>

struct c2g_nccl_api {
  bool ok = false;
  const char* err = "";
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclSend) Send = nullptr;
  decltype(&ncclRecv) Recv = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclBroadcast) Broadcast = nullptr;
};
c2g_nccl_api& c2g_nccl();  // loads on first use; check .ok

#define ncclGetUniqueId c2g_nccl().GetUniqueId
#define ncclCommInitRank c2g_nccl().CommInitRank
#define ncclCommDestroy c2g_nccl().CommDestroy
#define ncclGetErrorString c2g_nccl().GetErrorString
#define ncclGroupStart c2g_nccl().GroupStart
#define ncclGroupEnd c2g_nccl().GroupEnd
#define ncclSend c2g_nccl().Send
#define ncclRecv c2g_nccl().Recv
#define ncclAllReduce c2g_nccl().AllReduce
#define ncclAllGather c2g_nccl().AllGather
#define ncclBroadcast c2g_nccl().Broadcast
