// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// C wrapper around the JSON library THE REFERENCE VENDORS AND WRITES ITS STATE FILES WITH (src/nlohmann/json.hpp,
// nlohmann 3.4.0; `all_info.dump(4)` at src/NodeDataManager.cpp:619, src/PoseGraphSLAM.cpp:1199, src/Composer.cpp:1040),
// compiled from where it lies under /root/reference into oracle/_ref/libref_json.so by oracle/Makefile.
// tests/test_reference_json.py re-serialises the files the product writes through it: byte-identical output means the
// product's writer (csrc/host/Json.h) produces exactly what the reference's build would have produced for the same
// values — key order, indentation, number formatting — and that the reference's loader can read them.
#include <cstring>
#include <string>

#include "nlohmann/json.hpp"

using json = nlohmann::json;

static int copy_out(const std::string& s, char* out, int cap) {
  if (out && cap > 0) { const size_t n = s.size() < (size_t)(cap - 1) ? s.size() : (size_t)(cap - 1); std::memcpy(out, s.data(), n); out[n] = 0; }
  return (int)s.size();
}

extern "C" {

// parse `text`, dump(indent) it.  Returns the length of the dump (call with cap = 0 to size the buffer), -1 on a parse error.
int ref_json_redump(const char* text, int indent, char* out, int cap) {
  try { return copy_out(json::parse(text).dump(indent), out, cap); } catch (...) { return -1; }
}
// how the reference's library prints one double / one 64-bit integer
int ref_json_dump_double(double v, char* out, int cap) { try { return copy_out(json(v).dump(), out, cap); } catch (...) { return -1; } }
int ref_json_dump_int64(long long v, char* out, int cap) { try { return copy_out(json((std::int64_t)v).dump(), out, cap); } catch (...) { return -1; } }
// parse a number token the way the reference's loader would and hand back the double
int ref_json_parse_double(const char* token, double* out) {
  try { json j = json::parse(token); if (!j.is_number()) return 0; *out = j.get<double>(); return 1; } catch (...) { return 0; }
}

}  // extern "C"
