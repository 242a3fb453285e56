// Drop-in replacement for the reference's src/gcre_paths.h: class PathSet with the same public members
// (size, width_ul, vlen, operator[], set, load, select), but the rows live in B200 HBM behind the C ABI
// (include/gcre_b200.h).  Reference: src/gcre_paths.h:10-98.  Included by gcre.h after gcre_types.h, as upstream.
#ifndef GCRE_PATHS_H
#define GCRE_PATHS_H

#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../gcre_b200.h"
#include "gcre_types.h"

class PathSet;
using TPathSet = std::unique_ptr<PathSet>;

namespace gcre_detail {
// C status -> the reference's exception types (src/gcre_types.h:58-76, src/join_base.cpp:135)
inline void raise(int status) {
  if (status == GCRE_OK) return;
  if (status == GCRE_ERR_ASSERT) throw std::logic_error("assertion");
  if (status == GCRE_ERR_RANGE) throw std::out_of_range("assertion");
  throw std::runtime_error(std::string("gcre_b200: ") + gcre_last_error());
}
}  // namespace gcre_detail

// Run fn(device_slot) for every device slot; slots > 0 run on their own host threads (the C ABI is thread safe: one
// exec per device, thread-local error strings, a locked block cache).  Throws like gcre_detail::raise on the first failure.
template <typename Fn>
inline int gcre_for_each_device(size_t n_devices, const Fn& fn) {
  if (n_devices <= 1) {
    gcre_detail::raise(fn(0));  // throws the reference's exception types on failure
    return GCRE_OK;
  }
  std::vector<int> status(n_devices, GCRE_OK);
  std::vector<std::string> message(n_devices);
  std::vector<std::thread> pool;
  for (size_t d = 1; d < n_devices; d++)
    pool.emplace_back([&, d] {
      status[d] = fn(d);
      if (status[d] != GCRE_OK) message[d] = gcre_last_error();
    });
  status[0] = fn(0);
  if (status[0] != GCRE_OK) message[0] = gcre_last_error();
  for (auto& t : pool) t.join();
  for (size_t d = 0; d < n_devices; d++)
    if (status[d] != GCRE_OK) {
      if (status[d] == GCRE_ERR_ASSERT) throw std::logic_error("assertion");
      if (status[d] == GCRE_ERR_RANGE) throw std::out_of_range("assertion");
      throw std::runtime_error("gcre_b200: " + message[d]);
    }
  return GCRE_OK;
}

class PathSet {
 public:
  const st_pathset_size size;
  const uint16_t width_ul;  // words per half-row as the host sees them: ceil(n / 64)
  const uint16_t vlen;      // width_ul * method

  // takes ownership of device path sets created through the C ABI (JoinExec::createPathSet / select), one per GPU the
  // owning JoinExec drives (replicas hold identical rows)
  PathSet(std::vector<gcre_pathset*> handles, st_pathset_size size_, int width_ul_, uint16_t vlen_)
      : size(size_), width_ul((uint16_t)width_ul_), vlen(vlen_), handles_(std::move(handles)), row_cache_(vlen_) {}
  PathSet(const PathSet&) = delete;
  PathSet& operator=(const PathSet&) = delete;
  ~PathSet() {
    for (gcre_pathset* h : handles_) gcre_pathset_destroy(h);
  }

  // Row access (src/gcre_paths.h:44-47).  The row is fetched from the device into a per-object buffer; the pointer is
  // valid until the next operator[] on this object.
  const uint64_t* operator[](st_pathset_size idx) const {
    check_index(idx, size);
    gcre_detail::raise(gcre_pathset_get_row(handles_[0], idx, row_cache_.data()));
    return row_cache_.data();
  }

  // src/gcre_paths.h:49-52
  void set(st_pathset_size idx, const uint64_t* data) {
    check_index(idx, size);
    gcre_for_each_device(handles_.size(), [&](size_t d) { return gcre_pathset_set_row(handles_[d], idx, data); });
  }

  // src/gcre_paths.h:56-78: carriers (non-zero) go to the first half of each record; packed on the device
  void load(const vec2d_i& data) {
    check_true(size == data.size());
    const size_t cols = data.empty() ? 0 : data.front().size();
    std::vector<int32_t> flat(data.size() * cols);
    for (size_t r = 0; r < data.size(); r++) {
      check_equal(data[r].size(), cols);
      for (size_t c = 0; c < cols; c++) flat[r * cols + c] = data[r][c];
    }
    gcre_for_each_device(handles_.size(),
                         [&](size_t d) { return gcre_pathset_load_i32(handles_[d], flat.data(), (uint32_t)data.size(), (int)cols); });
  }

  // src/gcre_paths.h:82-92
  TPathSet select(const std::vector<int>& indices) const {
    std::vector<gcre_pathset*> out(handles_.size(), nullptr);
    gcre_for_each_device(handles_.size(),
                         [&](size_t d) { return gcre_pathset_select(handles_[d], indices.data(), (uint32_t)indices.size(), &out[d]); });
    return TPathSet(new PathSet(out, (st_pathset_size)indices.size(), width_ul, vlen));
  }

  gcre_pathset* handle(size_t device_slot = 0) const { return handles_[device_slot]; }

 protected:
  std::vector<gcre_pathset*> handles_;
  mutable std::vector<uint64_t> row_cache_;
};

#endif
