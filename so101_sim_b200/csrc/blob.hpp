// Host-side reader of the model blob written by tools/compile_model.py (format: so101_sim_b200/model.py).
#pragma once
#include <cstdint>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace so101 {

struct Blob {
  std::map<std::string, std::vector<double>> f;
  std::map<std::string, std::vector<int>> i;

  Blob(const void *data, size_t len) {
    const char *b = static_cast<const char *>(data);
    if (len < 16 || std::memcmp(b, "SO1B", 4) != 0) throw std::runtime_error("model blob: bad magic");
    uint32_t version, n;
    std::memcpy(&version, b + 4, 4); std::memcpy(&n, b + 8, 4);
    if (version != 2) throw std::runtime_error("model blob: unsupported version");
    for (uint32_t e = 0; e < n; e++) {
      const char *p = b + 16 + 40 * e;
      char name[25] = {0};
      std::memcpy(name, p, 24);
      uint32_t dt, cnt; uint64_t off;
      std::memcpy(&dt, p + 24, 4); std::memcpy(&cnt, p + 28, 4); std::memcpy(&off, p + 32, 8);
      if (off + (uint64_t)cnt * (dt == 0 ? 8 : 4) > len) throw std::runtime_error("model blob: truncated");
      if (dt == 0) { std::vector<double> v(cnt); std::memcpy(v.data(), b + off, cnt * 8); f[name] = std::move(v); }
      else { std::vector<int> v(cnt); std::memcpy(v.data(), b + off, cnt * 4); i[name] = std::move(v); }
    }
  }
  const std::vector<double> &F(const std::string &k) const {
    auto it = f.find(k);
    if (it == f.end()) throw std::runtime_error("model blob: missing float field " + k);
    return it->second;
  }
  const std::vector<int> &I(const std::string &k) const {
    auto it = i.find(k);
    if (it == i.end()) throw std::runtime_error("model blob: missing int field " + k);
    return it->second;
  }
  int scalar(const std::string &k) const { return I(k).at(0); }
};

}  // namespace so101
