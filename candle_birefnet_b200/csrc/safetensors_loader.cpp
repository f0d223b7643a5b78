// Minimal safetensors reader for brn_model_load_safetensors (include/birefnet_b200.h): replaces
// candle_core::safetensors::load + VarBuilder::from_tensors of examples/infer_image.rs:35-40.  Format (safetensors 0.x):
// u64 little-endian header length N, N bytes of JSON {"name": {"dtype": "F32", "shape": [..], "data_offsets": [a, b]},
// ..., "__metadata__": {...}}, then the raw little-endian tensor bytes.  Like VarBuilder, tensors the model does not
// ask for are ignored and a tensor the schema needs but the file lacks surfaces at finalize (BRN_ERR_MISSING_TENSOR).
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "model.h"

namespace brn {

namespace {
struct Cursor {
  const char* p; const char* e;
  void ws() { while (p < e && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p; }
  bool eat(char c) { ws(); if (p < e && *p == c) { ++p; return true; } return false; }
  void need(char c, const char* what) { if (!eat(c)) throw Error(5, std::string("safetensors header: expected '") + c + "' " + what); }
  std::string str() {
    ws();
    if (p >= e || *p != '"') throw Error(5, "safetensors header: expected a string");
    ++p;
    std::string s;
    while (p < e && *p != '"') {
      if (*p == '\\' && p + 1 < e) { ++p; s.push_back(*p == 'n' ? '\n' : *p == 't' ? '\t' : *p); ++p; }
      else s.push_back(*p++);
    }
    if (p >= e) throw Error(5, "safetensors header: unterminated string");
    ++p;
    return s;
  }
  long long num() {
    ws();
    long long v = 0; bool any = false;
    while (p < e && *p >= '0' && *p <= '9') { v = v * 10 + (*p - '0'); ++p; any = true; }
    if (!any) throw Error(5, "safetensors header: expected a number");
    return v;
  }
  void skip_value() {   // strings, numbers, nested objects / arrays (the __metadata__ entry)
    ws();
    if (p >= e) throw Error(5, "safetensors header: truncated");
    if (*p == '"') { str(); return; }
    if (*p == '{' || *p == '[') {
      const char open = *p, close = open == '{' ? '}' : ']';
      ++p;
      if (eat(close)) return;
      do {
        if (open == '{') { str(); need(':', "in object"); }
        skip_value();
      } while (eat(','));
      need(close, "closing a value");
      return;
    }
    while (p < e && *p != ',' && *p != '}' && *p != ']') ++p;
  }
};
}  // namespace

int load_safetensors(Model& m, const char* path) {
  FILE* f = fopen(path, "rb");
  BRN_CHECK(f != nullptr, 1, std::string("cannot open ") + path);
  std::vector<char> hdr;
  uint64_t n = 0;
  long data0 = 0;
  try {
    BRN_CHECK(fread(&n, 8, 1, f) == 1 && n > 1 && n < (1ull << 30), 5, "not a safetensors file (bad header length)");
    hdr.resize(n);
    BRN_CHECK(fread(hdr.data(), 1, n, f) == n, 5, "safetensors header truncated");
    data0 = 8 + (long)n;
    Cursor c{hdr.data(), hdr.data() + n};
    c.need('{', "at the start of the header");
    int loaded = 0;
    std::vector<char> buf;
    if (!c.eat('}')) {
      do {
        const std::string name = c.str();
        c.need(':', "after a tensor name");
        if (name == "__metadata__" || m.index.find(name) == m.index.end()) { c.skip_value(); continue; }   // VarBuilder ignores extras
        std::string dtype; std::vector<int64_t> shape; long long o0 = -1, o1 = -1;
        c.need('{', "opening a tensor entry");
        do {
          const std::string k = c.str();
          c.need(':', "in a tensor entry");
          if (k == "dtype") dtype = c.str();
          else if (k == "shape") { c.need('[', "shape"); if (!c.eat(']')) { do shape.push_back(c.num()); while (c.eat(',')); c.need(']', "shape"); } }
          else if (k == "data_offsets") { c.need('[', "data_offsets"); o0 = c.num(); c.need(',', "data_offsets"); o1 = c.num(); c.need(']', "data_offsets"); }
          else c.skip_value();
        } while (c.eat(','));
        c.need('}', "closing a tensor entry");
        int dt = dtype == "F32" ? BRN_F32 : dtype == "F16" ? BRN_F16 : dtype == "BF16" ? BRN_BF16 : -1;
        BRN_CHECK(dt >= 0, 5, "unsupported dtype " + dtype + " for tensor " + name);
        BRN_CHECK(o0 >= 0 && o1 >= o0, 5, "bad data_offsets for tensor " + name);
        size_t numel = 1;
        for (int64_t d : shape) numel *= (size_t)d;
        BRN_CHECK((size_t)(o1 - o0) == numel * (dt == BRN_F32 ? 4 : 2), 5, "byte size does not match the shape of " + name);
        buf.resize((size_t)(o1 - o0));
        BRN_CHECK(fseek(f, data0 + (long)o0, SEEK_SET) == 0 && fread(buf.data(), 1, buf.size(), f) == buf.size(), 5,
                  "safetensors data truncated at " + name);
        m.set_tensor(name.c_str(), buf.data(), dt, shape.data(), (int)shape.size());
        ++loaded;
      } while (c.eat(','));
      c.need('}', "at the end of the header");
    }
    fclose(f);
    return loaded;
  } catch (...) {
    fclose(f);
    throw;
  }
}

}  // namespace brn
