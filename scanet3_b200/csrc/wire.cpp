// Parameter wire format + checkpoint/resume.  Host-only code above the public C ABI
// (sb_param_info / sb_params_get / sb_params_set): no CUDA headers here.
//
// One tensor record is byte-for-byte what the reference's Kryo `TensorSerializer.write` emits
// (optimizers/KryoSerializers.scala:33-49) with Kryo 4.0.2 (Spark 3.3.1 -> chill 0.10.0), whose `Output.writeInt` /
// `writeLongs` are fixed-width BIG-endian and whose `writeBytes` is raw:
//     int32 BE  type code          (`TensorType.code` = TF DataType number; DT_FLOAT = 1, core/TensorType.scala:14,171)
//     int32 BE  rank
//     int64 BE  dims[rank]
//     int32 BE  payload bytes
//     payload   the tensor's buffer as is: row-major, NATIVE (little) endian fp32 (core/Tensor.scala:46-52)
// so a JVM maintainer reads a record with `TensorSerializer.read` and gets the `Tensor[Float]` of that path.
//
// A checkpoint file is this repo's own container around such records (the reference has no checkpointing, README.md:112):
//     "SB3CKPT1" | int64 BE global_iter | int64 BE epoch | int32 BE n | n x { int32 BE path bytes | path (UTF-8) | tensor record }
// holding every tensor of the arena: model weights, layer state and optimizer state, keyed by the reference's paths.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/scanet_b200.h"

extern "C" void sb_internal_set_last_error(const char* msg);  // engine.cu

namespace {

constexpr int32_t kDtFloat = 1;  // org.tensorflow.proto.framework.DataType.DT_FLOAT
constexpr char kMagic[8] = {'S', 'B', '3', 'C', 'K', 'P', 'T', '1'};

int fail(int code, const std::string& msg) {
  sb_internal_set_last_error(msg.c_str());
  return code;
}
void put32(uint8_t*& p, int32_t v) {
  for (int i = 3; i >= 0; --i) *p++ = (uint8_t)((uint32_t)v >> (8 * i));
}
void put64(uint8_t*& p, int64_t v) {
  for (int i = 7; i >= 0; --i) *p++ = (uint8_t)((uint64_t)v >> (8 * i));
}
int32_t get32(const uint8_t*& p) {
  uint32_t v = 0;
  for (int i = 0; i < 4; ++i) v = (v << 8) | *p++;
  return (int32_t)v;
}
int64_t get64(const uint8_t*& p) {
  uint64_t v = 0;
  for (int i = 0; i < 8; ++i) v = (v << 8) | *p++;
  return (int64_t)v;
}
size_t record_bytes(int rank, int64_t numel) { return 4 + 4 + 8 * (size_t)rank + 4 + 4 * (size_t)numel; }

struct Info {
  std::string path;
  int64_t dims[4] = {0, 0, 0, 0};
  int32_t rank = 0;
  int64_t numel = 1;
};
int info_of(const sb_engine* e, int i, Info* out) {
  char path[512];
  int32_t role, agg;
  int64_t off;
  int rc = sb_param_info(e, i, path, sizeof(path), out->dims, &out->rank, &role, &agg, &off);
  if (rc) return rc;
  out->path = path;
  out->numel = 1;
  for (int d = 0; d < out->rank; ++d) out->numel *= out->dims[d];
  return 0;
}
int find(const sb_engine* e, const char* path, Info* out) {
  if (!e || !path) return fail(SB_ERR_INVALID_ARG, "null argument");
  const int n = sb_param_count(e);
  for (int i = 0; i < n; ++i) {
    int rc = info_of(e, i, out);
    if (rc) return rc;
    if (out->path == path) return 0;
  }
  return fail(SB_ERR_INVALID_ARG, std::string("unknown parameter path: ") + path);
}

}  // namespace

extern "C" int sb_wire_encode(const int64_t* dims, int32_t rank, const float* data, void* buf, size_t cap, size_t* written) {
  if (rank < 0 || rank > 8 || (rank && !dims) || !written) return fail(SB_ERR_INVALID_ARG, "bad tensor description");
  int64_t numel = 1;
  for (int i = 0; i < rank; ++i) {
    if (dims[i] < 0) return fail(SB_ERR_INVALID_ARG, "negative dimension");
    numel *= dims[i];
  }
  if (numel * 4 > INT32_MAX) return fail(SB_ERR_UNSUPPORTED, "tensor above 2 GiB: TensorSerializer writes the payload size as Int");
  const size_t need = record_bytes(rank, numel);
  *written = need;
  if (!buf) return 0;  // size query
  if (cap < need) return fail(SB_ERR_INVALID_ARG, "wire buffer too small");
  if (!data && numel) return fail(SB_ERR_INVALID_ARG, "null data");
  uint8_t* p = (uint8_t*)buf;
  put32(p, kDtFloat);
  put32(p, rank);
  for (int i = 0; i < rank; ++i) put64(p, dims[i]);
  put32(p, (int32_t)(numel * 4));
  std::memcpy(p, data, (size_t)numel * 4);
  return 0;
}

extern "C" int sb_wire_decode(const void* buf, size_t n, int64_t* dims, int32_t dims_cap, int32_t* rank, float* data,
                              size_t data_cap, size_t* consumed) {
  if (!buf || !rank || !consumed) return fail(SB_ERR_INVALID_ARG, "null argument");
  const uint8_t* p = (const uint8_t*)buf;
  if (n < 8) return fail(SB_ERR_INVALID_ARG, "truncated tensor record");
  const int32_t code = get32(p);
  if (code != kDtFloat) return fail(SB_ERR_UNSUPPORTED, "only DT_FLOAT tensors cross this boundary (type code " + std::to_string(code) + ")");
  const int32_t r = get32(p);
  if (r < 0 || r > 8) return fail(SB_ERR_INVALID_ARG, "bad rank in tensor record");
  if (n < 8 + 8 * (size_t)r + 4) return fail(SB_ERR_INVALID_ARG, "truncated tensor record");
  int64_t numel = 1;
  for (int i = 0; i < r; ++i) {
    const int64_t d = get64(p);
    if (d < 0) return fail(SB_ERR_INVALID_ARG, "negative dimension in tensor record");
    if (dims && i < dims_cap) dims[i] = d;
    numel *= d;
  }
  const int32_t bytes = get32(p);
  if ((int64_t)bytes != numel * 4) return fail(SB_ERR_INVALID_ARG, "payload size does not match the shape");
  if (n < record_bytes(r, numel)) return fail(SB_ERR_INVALID_ARG, "truncated tensor payload");
  *rank = r;
  *consumed = record_bytes(r, numel);
  if (data) {
    if (dims && dims_cap < r) return fail(SB_ERR_INVALID_ARG, "dims buffer too small");
    if (data_cap < (size_t)numel) return fail(SB_ERR_INVALID_ARG, "data buffer too small");
    std::memcpy(data, p, (size_t)numel * 4);
  }
  return 0;
}

extern "C" int sb_param_to_wire(sb_engine* e, const char* path, void* buf, size_t cap, size_t* written) {
  Info in;
  int rc = find(e, path, &in);
  if (rc) return rc;
  if (!buf) return sb_wire_encode(in.dims, in.rank, nullptr, nullptr, 0, written);
  std::vector<float> host((size_t)in.numel);
  rc = sb_params_get(e, path, host.data(), host.size());
  if (rc) return rc;
  return sb_wire_encode(in.dims, in.rank, host.data(), buf, cap, written);
}

extern "C" int sb_param_from_wire(sb_engine* e, const char* path, const void* buf, size_t n, size_t* consumed) {
  Info in;
  int rc = find(e, path, &in);
  if (rc) return rc;
  int64_t dims[8] = {0};
  int32_t rank = 0;
  size_t used = 0;
  rc = sb_wire_decode(buf, n, dims, 8, &rank, nullptr, 0, &used);
  if (rc) return rc;
  bool same = rank == in.rank;
  for (int i = 0; same && i < rank; ++i) same = dims[i] == in.dims[i];
  if (!same) return fail(SB_ERR_INVALID_ARG, std::string("shape of the serialised tensor differs from parameter ") + path);
  std::vector<float> host((size_t)in.numel);
  rc = sb_wire_decode(buf, n, dims, 8, &rank, host.data(), host.size(), &used);
  if (rc) return rc;
  if (consumed) *consumed = used;
  return sb_params_set(e, path, host.data(), host.size());
}

extern "C" int sb_checkpoint_save(sb_engine* e, const char* file, int64_t global_iter, int64_t epoch) {
  if (!e || !file) return fail(SB_ERR_INVALID_ARG, "null argument");
  const int n = sb_param_count(e);
  const std::string tmp = std::string(file) + ".tmp";
  FILE* f = std::fopen(tmp.c_str(), "wb");
  if (!f) return fail(SB_ERR_STATE, "cannot open " + tmp + " for writing");
  auto bail = [&](int rc) { std::fclose(f); std::remove(tmp.c_str()); return rc; };
  uint8_t head[28];
  uint8_t* p = head;
  std::memcpy(p, kMagic, 8); p += 8;
  put64(p, global_iter);
  put64(p, epoch);
  put32(p, n);
  if (std::fwrite(head, 1, sizeof(head), f) != sizeof(head)) return bail(fail(SB_ERR_STATE, "short write"));
  std::vector<uint8_t> rec;
  for (int i = 0; i < n; ++i) {
    Info in;
    int rc = info_of(e, i, &in);
    if (rc) return bail(rc);
    size_t need = 0;
    rc = sb_param_to_wire(e, in.path.c_str(), nullptr, 0, &need);
    if (rc) return bail(rc);
    rec.resize(4 + in.path.size() + need);
    uint8_t* q = rec.data();
    put32(q, (int32_t)in.path.size());
    std::memcpy(q, in.path.data(), in.path.size());
    q += in.path.size();
    rc = sb_param_to_wire(e, in.path.c_str(), q, need, &need);
    if (rc) return bail(rc);
    if (std::fwrite(rec.data(), 1, rec.size(), f) != rec.size()) return bail(fail(SB_ERR_STATE, "short write"));
  }
  if (std::fclose(f) != 0) { std::remove(tmp.c_str()); return fail(SB_ERR_STATE, "close failed"); }
  if (std::rename(tmp.c_str(), file) != 0) { std::remove(tmp.c_str()); return fail(SB_ERR_STATE, "cannot move the checkpoint into place"); }
  return 0;
}

extern "C" int sb_internal_mark_params_ready(sb_engine* e);  // engine.cu: all tensors came from a checkpoint

extern "C" int sb_checkpoint_load(sb_engine* e, const char* file, int64_t* global_iter, int64_t* epoch) {
  if (!e || !file) return fail(SB_ERR_INVALID_ARG, "null argument");
  FILE* f = std::fopen(file, "rb");
  if (!f) return fail(SB_ERR_STATE, std::string("cannot open ") + file);
  std::vector<uint8_t> all;
  uint8_t chunk[1 << 16];
  size_t got;
  while ((got = std::fread(chunk, 1, sizeof(chunk), f)) > 0) all.insert(all.end(), chunk, chunk + got);
  std::fclose(f);
  if (all.size() < 28 || std::memcmp(all.data(), kMagic, 8) != 0) return fail(SB_ERR_INVALID_ARG, "not a scanet3_b200 checkpoint");
  const uint8_t* p = all.data() + 8;
  const uint8_t* end = all.data() + all.size();
  const int64_t it = get64(p);
  const int64_t ep = get64(p);
  const int32_t n = get32(p);
  if (n != sb_param_count(e)) return fail(SB_ERR_INVALID_ARG, "checkpoint holds a different number of tensors than this model + optimizer");
  for (int i = 0; i < n; ++i) {
    if (end - p < 4) return fail(SB_ERR_INVALID_ARG, "truncated checkpoint");
    const int32_t len = get32(p);
    if (len < 0 || end - p < len) return fail(SB_ERR_INVALID_ARG, "truncated checkpoint");
    const std::string path((const char*)p, (size_t)len);
    p += len;
    size_t used = 0;
    int rc = sb_param_from_wire(e, path.c_str(), p, (size_t)(end - p), &used);
    if (rc) return rc;
    p += used;
  }
  if (global_iter) *global_iter = it;
  if (epoch) *epoch = ep;
  return sb_internal_mark_params_ready(e);
}
