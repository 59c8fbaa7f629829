// swarm.cu -- osb_swarm: the swarm-wide keyframe exchange behind the C ABI.
//
// Replaces LoopNet::broadcast_fisheye_desc / image_desc_callback (swarm_loop/src/loop_net.cpp:20-120,142-172 of the
// reference: one LCM header message + one message per landmark over UDP multicast, reassembled with timeouts) by ONE
// ncclAllGather of the fixed-size osb_keyframe_record per keyframe round: on the 8-GPU box the drones are ranks of one
// communicator and the records move over NVLink.  The C++ nodelet (swarm_loop/src/swarm_loop.cpp:167) calls these entry
// points directly; nothing here needs Python.
//
// NCCL is opened with dlopen at the first osb_swarm_unique_id / osb_swarm_init, so libomniswarm_b200.so has no link-time
// dependency on it (a single-drone host without NCCL can still load the library; osb_swarm_* then return OSB_ERR_INVALID
// with a message).  world == 1 needs no NCCL at all.
//
// The reference's exchange is asynchronous (its LCM thread delivers remote keyframes whenever they arrive, loop_net.cpp
// :142-172), so nothing forces keyframe i's gather to finish before keyframe i+1's extraction starts:
// osb_swarm_exchange_async runs the collective on the handle's own stream behind an event of the caller's stream, and
// osb_swarm_wait makes a stream wait for it -- the caller ingests round i's foreign records while round i+1 is in flight.
#include <dlfcn.h>
#include <nccl.h>
#include "common.cuh"

namespace osb {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (!api.lib) return;
    api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.lib, "ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.lib, "ncclCommInitRank");
    api.AllGather = (decltype(api.AllGather))dlsym(api.lib, "ncclAllGather");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.lib, "ncclCommDestroy");
    api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.lib, "ncclGetErrorString");
    if (!api.GetUniqueId || !api.CommInitRank || !api.AllGather || !api.CommDestroy || !api.GetErrorString) api.lib = nullptr;
  });
  return api.lib ? &api : nullptr;
}

#define OSB_NCCL(api, expr)                                                                   \
  do {                                                                                        \
    ncclResult_t _r = (expr);                                                                 \
    if (_r != ncclSuccess) {                                                                  \
      char _buf[512];                                                                         \
      snprintf(_buf, sizeof(_buf), "%s:%d %s -> %s", __FILE__, __LINE__, #expr, (api)->GetErrorString(_r)); \
      osb::g_last_error = _buf;                                                               \
      return OSB_ERR_CUDA;                                                                    \
    }                                                                                         \
  } while (0)

}  // namespace osb

using namespace osb;

struct osb_swarm {
  int rank = 0, world = 1;
  ncclComm_t comm = nullptr;
  cudaStream_t side = nullptr;                 // the collective's own stream (exchange_async)
  cudaEvent_t ev_ready = nullptr, ev_done = nullptr;
  bool in_flight = false;
  std::mutex mu;
};

static_assert(sizeof(ncclUniqueId) == OSB_SWARM_ID_BYTES, "ncclUniqueId size");

extern "C" osb_status osb_swarm_unique_id(uint8_t* id_out) {
  OSB_REQUIRE(id_out != nullptr, "null id");
  NcclApi* api = nccl_api();
  if (!api) { set_error("osb_swarm_unique_id", "libnccl.so.2 could not be opened"); return OSB_ERR_INVALID; }
  ncclUniqueId id;
  OSB_NCCL(api, api->GetUniqueId(&id));
  memcpy(id_out, &id, sizeof(id));
  return OSB_OK;
}

extern "C" osb_status osb_swarm_destroy(osb_swarm* h) {
  if (!h) return OSB_OK;
  if (h->side) cudaStreamSynchronize(h->side);
  if (h->comm) { NcclApi* api = nccl_api(); if (api) api->CommDestroy(h->comm); }
  if (h->ev_ready) cudaEventDestroy(h->ev_ready);
  if (h->ev_done) cudaEventDestroy(h->ev_done);
  if (h->side) cudaStreamDestroy(h->side);
  delete h;
  return OSB_OK;
}

extern "C" osb_status osb_swarm_init(osb_swarm** out, const uint8_t* id, int rank, int world) {
  OSB_REQUIRE(out != nullptr && world >= 1 && rank >= 0 && rank < world, "bad rank / world");
  osb_status s = require_device();
  if (s != OSB_OK) return s;
  osb_swarm* h = new osb_swarm();
  h->rank = rank; h->world = world;
#define SW_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { set_error("osb_swarm_init", cudaGetErrorString(e_)); osb_swarm_destroy(h); return OSB_ERR_CUDA; } } while (0)
  SW_CUDA(cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking));
  SW_CUDA(cudaEventCreateWithFlags(&h->ev_ready, cudaEventDisableTiming));
  SW_CUDA(cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming));
#undef SW_CUDA
  if (world > 1) {
    NcclApi* api = nccl_api();
    if (!api || !id) {
      set_error("osb_swarm_init", api ? "null unique id" : "libnccl.so.2 could not be opened");
      osb_swarm_destroy(h);
      return OSB_ERR_INVALID;
    }
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    ncclResult_t r = api->CommInitRank(&h->comm, world, uid, rank);
    if (r != ncclSuccess) {
      set_error("osb_swarm_init: ncclCommInitRank", api->GetErrorString(r));
      h->comm = nullptr;
      osb_swarm_destroy(h);
      return OSB_ERR_CUDA;
    }
  }
  *out = h;
  return OSB_OK;
}

static osb_status swarm_gather(osb_swarm* h, const osb_keyframe_record* rec, osb_keyframe_record* gathered, cudaStream_t st) {
  if (h->world == 1) {
    if (gathered != rec) OSB_CUDA(cudaMemcpyAsync(gathered, rec, sizeof(osb_keyframe_record), cudaMemcpyDeviceToDevice, st));
    return OSB_OK;
  }
  NcclApi* api = nccl_api();
  OSB_NCCL(api, api->AllGather(rec, gathered, sizeof(osb_keyframe_record), ncclUint8, h->comm, st));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return OSB_OK;
}

extern "C" osb_status osb_swarm_exchange(osb_swarm* h, const osb_keyframe_record* record_dev,
                                         osb_keyframe_record* gathered_dev, void* stream) {
  OSB_REQUIRE(h && record_dev && gathered_dev, "null argument");
  std::lock_guard<std::mutex> lk(h->mu);
  return swarm_gather(h, record_dev, gathered_dev, (cudaStream_t)stream);
}

extern "C" osb_status osb_swarm_exchange_async(osb_swarm* h, const osb_keyframe_record* record_dev,
                                               osb_keyframe_record* gathered_dev, void* stream) {
  OSB_REQUIRE(h && record_dev && gathered_dev, "null argument");
  std::lock_guard<std::mutex> lk(h->mu);
  OSB_CUDA(cudaEventRecord(h->ev_ready, (cudaStream_t)stream));      // the record is complete on the caller's stream
  OSB_CUDA(cudaStreamWaitEvent(h->side, h->ev_ready, 0));
  osb_status s = swarm_gather(h, record_dev, gathered_dev, h->side);
  if (s != OSB_OK) return s;
  OSB_CUDA(cudaEventRecord(h->ev_done, h->side));
  h->in_flight = true;
  return OSB_OK;
}

extern "C" osb_status osb_swarm_wait(osb_swarm* h, void* stream) {
  OSB_REQUIRE(h != nullptr, "null handle");
  std::lock_guard<std::mutex> lk(h->mu);
  if (h->in_flight) OSB_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, h->ev_done, 0));
  return OSB_OK;
}

extern "C" int osb_swarm_rank(osb_swarm* h) { return h ? h->rank : -1; }
extern "C" int osb_swarm_world(osb_swarm* h) { return h ? h->world : -1; }
