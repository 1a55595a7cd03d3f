// Legacy-VTK unstructured-grid reader: the on-disk format on the way into the element path.
// The reference loads meshes with pyvista (`pv.read`, element.py:53) and takes `mesh.points` and the flat `mesh.cells`
// array `[nen, id0, .., nen, id0, ..]` (element.py:55-88).  This is a host-side parser of the same files:
//   # vtk DataFile Version x.y / title / ASCII|BINARY / DATASET UNSTRUCTURED_GRID
//   POINTS n <type>            coordinates (ASCII text or big-endian binary)
//   CELLS n size               classic layout: per cell `nen id..`  (versions <= 4.2)
//   CELLS n+1 m + OFFSETS <type> + CONNECTIVITY <type>   (version 5.x)
//   CELL_TYPES n
// Everything after CELL_TYPES (point / cell data) is ignored.  Host code only -- nothing here touches the GPU.
#include <cerrno>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "common.cuh"

struct femb_vtk_mesh {
  std::vector<double> points;     // [n,3]
  std::vector<long long> cells;   // legacy flat layout
  std::vector<int> types;         // [n_cells] (may be empty if the file ends before CELL_TYPES)
  long long n_points = 0, n_cells = 0;
  bool points_are_float = false;  // file stores 32-bit coordinates (pyvista then returns float32 points)
};

namespace femb {
namespace {

struct Reader {
  std::string buf;
  size_t pos = 0;
  bool binary = false;
  std::string err;

  bool eof() const { return pos >= buf.size(); }
  void skip_ws() {
    while (pos < buf.size() && (buf[pos] == ' ' || buf[pos] == '\n' || buf[pos] == '\r' || buf[pos] == '\t')) ++pos;
  }
  std::string line() {
    size_t e = buf.find('\n', pos);
    if (e == std::string::npos) e = buf.size();
    std::string s = buf.substr(pos, e - pos);
    pos = e < buf.size() ? e + 1 : e;
    if (!s.empty() && s.back() == '\r') s.pop_back();
    return s;
  }
  std::string token() {
    skip_ws();
    const size_t b = pos;
    while (pos < buf.size() && !(buf[pos] == ' ' || buf[pos] == '\n' || buf[pos] == '\r' || buf[pos] == '\t')) ++pos;
    return buf.substr(b, pos - b);
  }
  // after a header line of a BINARY section the payload starts right after the single newline
  void to_payload() {
    while (pos < buf.size() && (buf[pos] == ' ' || buf[pos] == '\r' || buf[pos] == '\t')) ++pos;
    if (pos < buf.size() && buf[pos] == '\n') ++pos;
  }
  static int type_bytes(const std::string& t, bool* is_float, bool* is_unsigned) {
    *is_float = false, *is_unsigned = false;
    if (t == "float") { *is_float = true; return 4; }
    if (t == "double") { *is_float = true; return 8; }
    if (t == "char") return 1;
    if (t == "unsigned_char") { *is_unsigned = true; return 1; }
    if (t == "short") return 2;
    if (t == "unsigned_short") { *is_unsigned = true; return 2; }
    if (t == "int" || t == "vtktypeint32") return 4;
    if (t == "unsigned_int" || t == "vtktypeuint32") { *is_unsigned = true; return 4; }
    if (t == "long" || t == "vtktypeint64" || t == "vtkidtype") return 8;
    if (t == "unsigned_long" || t == "vtktypeuint64") { *is_unsigned = true; return 8; }
    return 0;
  }
  // big-endian payload -> host values
  template <typename OUT>
  bool read_binary(const std::string& type, long long count, OUT* out) {
    bool isf, isu;
    const int nb = type_bytes(type, &isf, &isu);
    if (!nb) { err = "unsupported data type '" + type + "'"; return false; }
    if (pos + (size_t)nb * count > buf.size()) { err = "file truncated inside a binary section"; return false; }
    const unsigned char* p = reinterpret_cast<const unsigned char*>(buf.data()) + pos;
    for (long long k = 0; k < count; ++k, p += nb) {
      unsigned long long v = 0;
      for (int b = 0; b < nb; ++b) v = (v << 8) | p[b];
      if (isf) {
        if (nb == 4) { unsigned int w = (unsigned int)v; float f; memcpy(&f, &w, 4); out[k] = (OUT)f; }
        else { double d; memcpy(&d, &v, 8); out[k] = (OUT)d; }
      } else if (isu) {
        out[k] = (OUT)v;
      } else {
        long long sv = nb == 8 ? (long long)v : nb == 4 ? (long long)(int)(unsigned int)v : nb == 2 ? (long long)(short)(unsigned short)v
                                                                                                     : (long long)(signed char)(unsigned char)v;
        out[k] = (OUT)sv;
      }
    }
    pos += (size_t)nb * count;
    return true;
  }
  bool read_ascii_real(long long count, bool as_float, double* out) {
    for (long long k = 0; k < count; ++k) {
      skip_ws();
      if (eof()) { err = "file truncated inside POINTS"; return false; }
      char* end = nullptr;
      const char* b = buf.c_str() + pos;
      // a `float` file is parsed to float32 exactly as VTK does, so the later cast to the requested dtype sees the same bits
      out[k] = as_float ? (double)strtof(b, &end) : strtod(b, &end);
      if (end == b) { err = "non-numeric token inside POINTS"; return false; }
      pos += (size_t)(end - b);
    }
    return true;
  }
  bool read_ascii_int(long long count, long long* out, const char* where) {
    for (long long k = 0; k < count; ++k) {
      skip_ws();
      if (eof()) { err = std::string("file truncated inside ") + where; return false; }
      char* end = nullptr;
      const char* b = buf.c_str() + pos;
      out[k] = strtoll(b, &end, 10);
      if (end == b) { err = std::string("non-integer token inside ") + where; return false; }
      pos += (size_t)(end - b);
    }
    return true;
  }
};

std::string upper(std::string s) {
  for (auto& c : s) c = (char)toupper((unsigned char)c);
  return s;
}
std::string lower(std::string s) {
  for (auto& c : s) c = (char)tolower((unsigned char)c);
  return s;
}

int parse_vtk(const char* path, femb_vtk_mesh* m) {
  Reader rd;
  {
    std::ifstream f(path, std::ios::binary);
    if (!f) { set_error(std::string("femb_vtk_open: cannot open '") + path + "'"); return FEMB_ERR_ARG; }
    std::ostringstream ss;
    ss << f.rdbuf();
    rd.buf = ss.str();
  }
  const std::string magic = rd.line();
  if (magic.compare(0, 5, "# vtk") != 0) { set_error("femb_vtk_open: not a legacy VTK file (missing '# vtk DataFile' header); XML .vtu files are not supported"); return FEMB_ERR_UNSUPPORTED; }
  rd.line();  // title
  const std::string fmt = upper(rd.token());
  if (fmt == "BINARY") rd.binary = true;
  else if (fmt != "ASCII") { set_error("femb_vtk_open: expected ASCII or BINARY, got '" + fmt + "'"); return FEMB_ERR_ARG; }
  bool have_points = false, have_cells = false;
  while (true) {
    const std::string key = upper(rd.token());
    if (key.empty()) break;
    if (key == "DATASET") {
      const std::string kind = upper(rd.token());
      if (kind != "UNSTRUCTURED_GRID") { set_error("femb_vtk_open: DATASET " + kind + " is not supported (UNSTRUCTURED_GRID only)"); return FEMB_ERR_UNSUPPORTED; }
    } else if (key == "FIELD") {  // e.g. the FieldData block some writers put before POINTS: name n ; then arrays `name comps tuples type`
      rd.token();
      const long long narr = strtoll(rd.token().c_str(), nullptr, 10);
      for (long long a = 0; a < narr; ++a) {
        rd.token();
        const long long comps = strtoll(rd.token().c_str(), nullptr, 10), tuples = strtoll(rd.token().c_str(), nullptr, 10);
        const std::string type = lower(rd.token());
        if (rd.binary) {
          bool isf, isu;
          const int nb = Reader::type_bytes(type, &isf, &isu);
          if (!nb) { set_error("femb_vtk_open: unsupported FIELD array type '" + type + "'"); return FEMB_ERR_UNSUPPORTED; }
          rd.to_payload();
          rd.pos += (size_t)nb * comps * tuples;
        } else {
          for (long long k = 0; k < comps * tuples; ++k) rd.token();
        }
      }
    } else if (key == "POINTS") {
      m->n_points = strtoll(rd.token().c_str(), nullptr, 10);
      const std::string type = lower(rd.token());
      if (m->n_points < 0) { set_error("femb_vtk_open: negative POINTS count"); return FEMB_ERR_ARG; }
      m->points_are_float = type == "float";
      m->points.resize((size_t)m->n_points * 3);
      bool ok;
      if (rd.binary) {
        rd.to_payload();
        ok = rd.read_binary<double>(type, m->n_points * 3, m->points.data());
      } else {
        ok = rd.read_ascii_real(m->n_points * 3, m->points_are_float, m->points.data());
      }
      if (!ok) { set_error("femb_vtk_open: " + rd.err); return FEMB_ERR_ARG; }
      have_points = true;
    } else if (key == "CELLS") {
      const long long a = strtoll(rd.token().c_str(), nullptr, 10), b = strtoll(rd.token().c_str(), nullptr, 10);
      if (a < 0 || b < 0) { set_error("femb_vtk_open: negative CELLS sizes"); return FEMB_ERR_ARG; }
      // version 5.x writes `CELLS n+1 m` followed by OFFSETS / CONNECTIVITY; classic files continue with numbers
      const size_t save = rd.pos;
      const std::string next = upper(rd.token());
      bool ok = true;
      if (next == "OFFSETS") {
        const std::string otype = lower(rd.token());
        std::vector<long long> off((size_t)a), con((size_t)b);
        if (rd.binary) { rd.to_payload(); ok = rd.read_binary<long long>(otype, a, off.data()); }
        else ok = rd.read_ascii_int(a, off.data(), "OFFSETS");
        if (ok) {
          const std::string ckey = upper(rd.token());
          if (ckey != "CONNECTIVITY") { set_error("femb_vtk_open: expected CONNECTIVITY after OFFSETS"); return FEMB_ERR_ARG; }
          const std::string ctype = lower(rd.token());
          if (rd.binary) { rd.to_payload(); ok = rd.read_binary<long long>(ctype, b, con.data()); }
          else ok = rd.read_ascii_int(b, con.data(), "CONNECTIVITY");
        }
        if (ok) {
          m->n_cells = a > 0 ? a - 1 : 0;
          m->cells.clear();
          m->cells.reserve((size_t)(m->n_cells + b));
          for (long long c = 0; c < m->n_cells; ++c) {
            const long long s0 = off[c], s1 = off[c + 1];
            if (s0 < 0 || s1 < s0 || s1 > b) { set_error("femb_vtk_open: OFFSETS are not monotone / exceed CONNECTIVITY"); return FEMB_ERR_ARG; }
            m->cells.push_back(s1 - s0);
            for (long long k = s0; k < s1; ++k) m->cells.push_back(con[k]);
          }
        }
      } else {
        rd.pos = save;
        m->n_cells = a;
        m->cells.resize((size_t)b);
        if (rd.binary) { rd.to_payload(); ok = rd.read_binary<long long>("int", b, m->cells.data()); }
        else ok = rd.read_ascii_int(b, m->cells.data(), "CELLS");
        if (ok) {  // validate the `nen id..` structure
          long long k = 0, c = 0;
          for (; c < a && k < b; ++c) {
            if (m->cells[k] < 0) break;
            k += 1 + m->cells[k];
          }
          if (c != a || k != b) { set_error("femb_vtk_open: CELLS size does not match its per-cell counts"); return FEMB_ERR_ARG; }
        }
      }
      if (!ok) { set_error("femb_vtk_open: " + rd.err); return FEMB_ERR_ARG; }
      have_cells = true;
    } else if (key == "CELL_TYPES") {
      const long long n = strtoll(rd.token().c_str(), nullptr, 10);
      if (n < 0) { set_error("femb_vtk_open: negative CELL_TYPES count"); return FEMB_ERR_ARG; }
      std::vector<long long> t((size_t)n);
      bool ok;
      if (rd.binary) { rd.to_payload(); ok = rd.read_binary<long long>("int", n, t.data()); }
      else ok = rd.read_ascii_int(n, t.data(), "CELL_TYPES");
      if (!ok) { set_error("femb_vtk_open: " + rd.err); return FEMB_ERR_ARG; }
      m->types.assign(t.begin(), t.end());
      break;  // attribute data follows; not needed
    } else if (key == "POINT_DATA" || key == "CELL_DATA") {
      break;
    } else if (key == "METADATA") {  // skipped up to the blank line that ends it
      while (!rd.eof() && !rd.line().empty()) {}
    } else {
      set_error("femb_vtk_open: unexpected keyword '" + key + "'");
      return FEMB_ERR_ARG;
    }
  }
  if (!have_points || !have_cells) { set_error("femb_vtk_open: file has no POINTS / CELLS section"); return FEMB_ERR_ARG; }
  for (size_t k = 0, c = 0; c < (size_t)m->n_cells; ++c) {  // ids must address the point array
    const long long nen = m->cells[k++];
    for (long long j = 0; j < nen; ++j, ++k)
      if (m->cells[k] < 0 || m->cells[k] >= m->n_points) { set_error("femb_vtk_open: cell references a point id outside POINTS"); return FEMB_ERR_ARG; }
  }
  return FEMB_OK;
}

}  // namespace
}  // namespace femb

using namespace femb;

extern "C" int femb_vtk_open(const char* path, femb_vtk_mesh** mesh, int64_t* n_points, int64_t* n_cells, int64_t* cells_size,
                             int32_t* points_are_float) {
  FEMB_CHECK_ARG(path && mesh, "path / mesh");
  auto* m = new femb_vtk_mesh();
  const int rc = parse_vtk(path, m);
  if (rc != FEMB_OK) {
    delete m;
    return rc;
  }
  *mesh = m;
  if (n_points) *n_points = m->n_points;
  if (n_cells) *n_cells = m->n_cells;
  if (cells_size) *cells_size = (int64_t)m->cells.size();
  if (points_are_float) *points_are_float = m->points_are_float ? 1 : 0;
  return FEMB_OK;
}

extern "C" int femb_vtk_read(femb_vtk_mesh* m, double* points_host, int64_t* cells_host, int32_t* types_host) {
  FEMB_CHECK_ARG(m != nullptr, "mesh");
  if (points_host && !m->points.empty()) memcpy(points_host, m->points.data(), sizeof(double) * m->points.size());
  if (cells_host)
    for (size_t k = 0; k < m->cells.size(); ++k) cells_host[k] = m->cells[k];
  if (types_host) {
    FEMB_CHECK_ARG((long long)m->types.size() == m->n_cells, "the file has no (complete) CELL_TYPES section");
    for (size_t k = 0; k < m->types.size(); ++k) types_host[k] = m->types[k];
  }
  return FEMB_OK;
}

extern "C" int femb_vtk_close(femb_vtk_mesh* m) {
  delete m;
  return FEMB_OK;
}
