// m3b_file.cu -- the spline monolith on disk without ROOT.
//
// The reference caches a built SMonolith in a ROOT file (SMonolith::PrepareSplineFile / LoadSplineFile,
// Splines/SplineMonolith.cpp:454-614: trees "Settings", "SplineTree/SplineObject" = the serialised SplineMonoStruct,
// "Monolith_TF1", "EventInfo", and the FastSplineInfo directory of Splines/SplineBase.cpp:139-191) so that the
// expensive per-event TSpline3 flattening runs once.  ROOT is not a dependency of this library; the same arrays are
// written as ONE flat little-endian file whose sections carry the reference's member names:
//
//     Header   magic "M3BMONO1", version, n_sections, and the scalars of the "Settings" tree
//              (NEvents, nParams, _max_knots, nKnots, NSplines_valid, NTF1_valid)
//     Section  table {name[24], bytes per element, count, file offset}, data 64-byte aligned:
//              coeff_x, coeff_many, nKnots_arr, paramNo_arr                (SplineMonoStruct, Splines/SplineCommon.h:30-50)
//              cpu_nParamPerEvent, cpu_nParamPerEvent_tf1                  ("EventInfo")
//              cpu_coeff_TF1_many, cpu_paramNo_TF1_arr                     ("Monolith_TF1")
//              nPts, xPts_f64 (optional)                                   (FastSplineInfo)
//
//   m3b_write_monolith_file   what a MaCh3 maintainer calls next to PrepareSplineFile (adapters/M3BMonolithFile.h)
//   m3b_upload_from_file      streams the file to the device in chunks of events (bounded host memory: a config-3
//                             monolith is 109 GB on disk, ~0.8 GB resident while loading) through m3b_splines_append
//   m3b_group_upload_from_file every member of a single-process group reads its own event range
#include "m3b_handle.h"

#include <cstdio>
#include <unistd.h>

namespace {

constexpr char kMagic[8] = {'M', '3', 'B', 'M', 'O', 'N', 'O', '1'};

struct FileHeader {
  char magic[8];
  uint32_t version, n_sections;
  uint64_t n_events;
  int32_t n_params, max_knots;
  uint64_t n_splines_valid, n_tf1_valid, total_knots;
  uint32_t has_xpts, reserved;
};
struct FileSection {
  char name[24];
  uint32_t elem_bytes, reserved;
  uint64_t count, offset;
};
static_assert(sizeof(FileHeader) == 64 && sizeof(FileSection) == 48, "file structures are packed as documented");

struct Reader {
  FILE* f = nullptr;
  FileHeader hd{};
  std::vector<FileSection> sec;
  ~Reader() { if (f) fclose(f); }
  const FileSection* find(const char* name) const {
    for (const FileSection& s : sec) if (!strncmp(s.name, name, sizeof s.name)) return &s;
    return nullptr;
  }
  bool read(const FileSection* s, uint64_t first, uint64_t n, void* out) const {
    if (!s || first + n > s->count) return false;
    if (n == 0) return true;
    if (fseeko(f, static_cast<off_t>(s->offset + first * s->elem_bytes), SEEK_SET) != 0) return false;
    return fread(out, s->elem_bytes, n, f) == n;
  }
};

int open_reader(m3b_handle* h, const char* path, Reader& r) {
  REQUIRE(path, M3B_ERR_INVALID, "monolith file: null path");
  r.f = fopen(path, "rb");
  REQUIRE(r.f, M3B_ERR_INVALID, std::string("monolith file: cannot open ") + path);
  REQUIRE(fread(&r.hd, sizeof r.hd, 1, r.f) == 1 && !memcmp(r.hd.magic, kMagic, 8), M3B_ERR_INVALID,
          std::string("monolith file: ") + path + " is not an M3BMONO1 file");
  REQUIRE(r.hd.version == 1 && r.hd.n_sections <= 64, M3B_ERR_INVALID, "monolith file: unsupported version");
  r.sec.resize(r.hd.n_sections);
  REQUIRE(fread(r.sec.data(), sizeof(FileSection), r.sec.size(), r.f) == r.sec.size(), M3B_ERR_INVALID, "monolith file: truncated section table");
  for (const char* need : {"coeff_x", "coeff_many", "nKnots_arr", "paramNo_arr", "cpu_nParamPerEvent", "cpu_nParamPerEvent_tf1",
                           "cpu_coeff_TF1_many", "cpu_paramNo_TF1_arr", "nPts"})
    REQUIRE(r.find(need), M3B_ERR_INVALID, std::string("monolith file: section missing: ") + need);
  return M3B_OK;
}

// events [e0, e1) of the file into handle h (its whole monolith), in chunks
int upload_range(m3b_handle* h, const Reader& r, int64_t e0, int64_t e1, int64_t chunk_events) {
  const FileHeader& hd = r.hd;
  const int P = hd.n_params, K = hd.max_knots;
  std::vector<float> coeff_x(static_cast<size_t>(P) * K);
  std::vector<int16_t> n_pts(P);
  REQUIRE(r.read(r.find("coeff_x"), 0, coeff_x.size(), coeff_x.data()) && r.read(r.find("nPts"), 0, n_pts.size(), n_pts.data()),
          M3B_ERR_INVALID, "monolith file: cannot read the knot tables");
  int rc = m3b_splines_begin(h, P, K, coeff_x.data(), n_pts.data(), e1 - e0);
  if (rc != M3B_OK) return rc;
  // response offsets of the first event of the range: prefix sums over the {count,start} pairs before it
  const FileSection* s_npe = r.find("cpu_nParamPerEvent");
  const FileSection* s_npl = r.find("cpu_nParamPerEvent_tf1");
  uint64_t oc = 0, ol = 0;
  {
    std::vector<uint32_t> blk;
    for (int64_t b0 = 0; b0 < e0; b0 += 1 << 20) {
      const int64_t n = std::min<int64_t>(1 << 20, e0 - b0);
      blk.resize(2 * n);
      REQUIRE(r.read(s_npe, 2 * b0, 2 * n, blk.data()), M3B_ERR_INVALID, "monolith file: truncated cpu_nParamPerEvent");
      for (int64_t i = 0; i < n; ++i) oc += blk[2 * i];
      REQUIRE(r.read(s_npl, 2 * b0, 2 * n, blk.data()), M3B_ERR_INVALID, "monolith file: truncated cpu_nParamPerEvent_tf1");
      for (int64_t i = 0; i < n; ++i) ol += blk[2 * i];
    }
  }
  if (chunk_events <= 0) chunk_events = 131072;
  chunk_events = std::max<int64_t>(h->T, chunk_events / h->T * h->T);       // every chunk but the last holds whole tile rows
  std::vector<uint32_t> npe, npl, koff32;
  std::vector<int16_t> pno, pnl;
  std::vector<uint64_t> koff;
  std::vector<float> many, tf1;
  const FileSection* s_k = r.find("nKnots_arr");
  for (int64_t c0 = e0; c0 < e1; c0 += chunk_events) {
    const int64_t n = std::min<int64_t>(chunk_events, e1 - c0);
    npe.resize(2 * n); npl.resize(2 * n);
    REQUIRE(r.read(s_npe, 2 * c0, 2 * n, npe.data()) && r.read(s_npl, 2 * c0, 2 * n, npl.data()), M3B_ERR_INVALID, "monolith file: truncated event table");
    uint64_t nc = 0, nl = 0;
    for (int64_t i = 0; i < n; ++i) { nc += npe[2 * i]; nl += npl[2 * i]; }
    pno.resize(nc); koff32.resize(nc + 1); pnl.resize(nl); tf1.resize(2 * nl);
    REQUIRE(r.read(r.find("paramNo_arr"), oc, nc, pno.data()), M3B_ERR_INVALID, "monolith file: truncated paramNo_arr");
    // first-knot offsets of the chunk's responses, and of the response after the chunk (or the end of the knots)
    const bool more = oc + nc < s_k->count;
    REQUIRE(r.read(s_k, oc, nc + (more ? 1 : 0), koff32.data()), M3B_ERR_INVALID, "monolith file: truncated nKnots_arr");
    const uint64_t k0 = nc ? koff32[0] : 0, k1 = nc ? (more ? koff32[nc] : hd.total_knots) : 0;
    REQUIRE(k1 >= k0 && k1 <= hd.total_knots, M3B_ERR_INVALID, "monolith file: nKnots_arr is not increasing");
    koff.resize(nc);
    for (uint64_t s = 0; s < nc; ++s) koff[s] = koff32[s] - k0;
    many.resize(4 * (k1 - k0));
    REQUIRE(r.read(r.find("coeff_many"), 4 * k0, 4 * (k1 - k0), many.data()), M3B_ERR_INVALID, "monolith file: truncated coeff_many");
    REQUIRE(r.read(r.find("cpu_paramNo_TF1_arr"), ol, nl, pnl.data()) && r.read(r.find("cpu_coeff_TF1_many"), 2 * ol, 2 * nl, tf1.data()),
            M3B_ERR_INVALID, "monolith file: truncated TF1 arrays");
    rc = m3b_splines_append(h, n, npe.data(), pno.data(), koff.data(), k1 - k0, many.data(), npl.data(), pnl.data(), tf1.data());
    if (rc != M3B_OK) return rc;
    oc += nc; ol += nl;
  }
  rc = m3b_splines_end(h);
  if (rc != M3B_OK) return rc;
  if (hd.has_xpts) {
    std::vector<double> x(static_cast<size_t>(P) * K);
    REQUIRE(r.read(r.find("xPts_f64"), 0, x.size(), x.data()), M3B_ERR_INVALID, "monolith file: truncated xPts_f64");
    rc = m3b_set_spline_knots_f64(h, x.data());
  }
  return rc;
}

}  // namespace

extern "C" {

M3B_API int m3b_write_monolith_file(const char* path, int32_t n_params, int32_t max_knots, const float* coeff_x,
                                    const int16_t* n_pts, const double* x_pts_f64, int64_t n_events,
                                    const uint32_t* nParamPerEvent, const int16_t* paramNo_arr, const uint32_t* nKnots_arr,
                                    uint32_t total_knots, const float* coeff_many, const uint32_t* nParamPerEvent_tf1,
                                    const int16_t* paramNo_tf1, const float* coeff_tf1) {
  m3b_handle* h = nullptr;
  REQUIRE(path && coeff_x && n_pts && nParamPerEvent && nParamPerEvent_tf1 && n_events >= 0 && n_params > 0 && max_knots >= 0,
          M3B_ERR_INVALID, "m3b_write_monolith_file: bad argument");
  uint64_t ns = 0, nl = 0;
  for (int64_t e = 0; e < n_events; ++e) { ns += nParamPerEvent[2 * e]; nl += nParamPerEvent_tf1[2 * e]; }
  REQUIRE(ns == 0 || (paramNo_arr && nKnots_arr && coeff_many), M3B_ERR_INVALID, "m3b_write_monolith_file: null TSpline3 arrays");
  REQUIRE(nl == 0 || (paramNo_tf1 && coeff_tf1), M3B_ERR_INVALID, "m3b_write_monolith_file: null TF1 arrays");
  struct Item { const char* name; uint32_t bytes; uint64_t count; const void* data; };
  std::vector<Item> items = {
      {"coeff_x", 4, static_cast<uint64_t>(n_params) * max_knots, coeff_x},
      {"coeff_many", 4, 4ull * total_knots, coeff_many},
      {"nKnots_arr", 4, ns, nKnots_arr},
      {"paramNo_arr", 2, ns, paramNo_arr},
      {"cpu_nParamPerEvent", 4, 2ull * n_events, nParamPerEvent},
      {"cpu_nParamPerEvent_tf1", 4, 2ull * n_events, nParamPerEvent_tf1},
      {"cpu_coeff_TF1_many", 4, 2 * nl, coeff_tf1},
      {"cpu_paramNo_TF1_arr", 2, nl, paramNo_tf1},
      {"nPts", 2, static_cast<uint64_t>(n_params), n_pts}};
  if (x_pts_f64) items.push_back({"xPts_f64", 8, static_cast<uint64_t>(n_params) * max_knots, x_pts_f64});
  FileHeader hd{};
  memcpy(hd.magic, kMagic, 8);
  hd.version = 1; hd.n_sections = static_cast<uint32_t>(items.size());
  hd.n_events = static_cast<uint64_t>(n_events); hd.n_params = n_params; hd.max_knots = max_knots;
  hd.n_splines_valid = ns; hd.n_tf1_valid = nl; hd.total_knots = total_knots; hd.has_xpts = x_pts_f64 ? 1 : 0;
  std::vector<FileSection> sec(items.size());
  uint64_t off = (sizeof hd + sizeof(FileSection) * items.size() + 63) & ~63ull;
  for (size_t i = 0; i < items.size(); ++i) {
    memset(&sec[i], 0, sizeof sec[i]);
    strncpy(sec[i].name, items[i].name, sizeof sec[i].name - 1);
    sec[i].elem_bytes = items[i].bytes; sec[i].count = items[i].count; sec[i].offset = off;
    off = (off + items[i].bytes * items[i].count + 63) & ~63ull;
  }
  FILE* f = fopen(path, "wb");
  REQUIRE(f, M3B_ERR_INVALID, std::string("m3b_write_monolith_file: cannot create ") + path);
  bool ok = fwrite(&hd, sizeof hd, 1, f) == 1 && fwrite(sec.data(), sizeof(FileSection), sec.size(), f) == sec.size();
  for (size_t i = 0; ok && i < items.size(); ++i) {
    ok = fseeko(f, static_cast<off_t>(sec[i].offset), SEEK_SET) == 0;
    if (ok && items[i].count) ok = fwrite(items[i].data, items[i].bytes, items[i].count, f) == items[i].count;
  }
  if (ok) ok = fflush(f) == 0 && ftruncate(fileno(f), static_cast<off_t>(off)) == 0;      // extend to the aligned end
  ok = (fclose(f) == 0) && ok;
  REQUIRE(ok, M3B_ERR_INVALID, std::string("m3b_write_monolith_file: write failed: ") + path);
  return M3B_OK;
}

M3B_API int m3b_monolith_file_info(const char* path, int64_t* n_events, int32_t* n_params, int32_t* max_knots, uint64_t* total_knots) {
  m3b_handle* h = nullptr;
  Reader r;
  int rc = open_reader(h, path, r);
  if (rc != M3B_OK) return rc;
  if (n_events) *n_events = static_cast<int64_t>(r.hd.n_events);
  if (n_params) *n_params = r.hd.n_params;
  if (max_knots) *max_knots = r.hd.max_knots;
  if (total_knots) *total_knots = r.hd.total_knots;
  return M3B_OK;
}

M3B_API int m3b_upload_from_file(m3b_handle* h, const char* path, int64_t chunk_events) {
  REQUIRE(h, M3B_ERR_INVALID, "null handle");
  Reader r;
  int rc = open_reader(h, path, r);
  if (rc != M3B_OK) return rc;
  return upload_range(h, r, 0, static_cast<int64_t>(r.hd.n_events), chunk_events);
}

M3B_API int m3b_group_upload_from_file(m3b_group* g, const char* path, int64_t chunk_events) {
  m3b_handle* h = m3b_group_member(g, 0);
  REQUIRE(h, M3B_ERR_INVALID, "m3b_group_upload_from_file: null group");
  Reader r;
  int rc = open_reader(h, path, r);
  if (rc != M3B_OK) return rc;
  for (int32_t i = 0; i < m3b_group_size(g); ++i) {
    int64_t e0 = 0, e1 = 0;
    m3b_group_shard(g, static_cast<int64_t>(r.hd.n_events), i, &e0, &e1);
    rc = upload_range(m3b_group_member(g, i), r, e0, e1, chunk_events);
    if (rc != M3B_OK) return rc;
  }
  return M3B_OK;
}

}  // extern "C"
