// Flat, mmap-able snapshot of the trained object DB: what DescriptorMatcher::parameter_callback
// (src/detection/DescriptorMatcher.cpp:60-129 of the reference) pulls out of CouchDB document by document — field
// "object_id" (:72), attachments "descriptors" N x 32 CV_8U (:74-76, written by training.cpp:157 / ModelFiller.cpp:23)
// and "points" 1 x N CV_32FC3 (:79-86, training.cpp:158 / ModelFiller.cpp:24) — in one file that is read with a single
// mmap, no DB server and no cv::Mat deserialisation.  Host-only: none of these entry points needs a GPU.
//
// Layout (little endian, every section 64-byte aligned):
//   header   64 B   magic "TODB200\0", u32 version = 1, u32 n_objects, u64 total_rows,
//                   u64 off_table, u64 off_descriptors, u64 off_points, u64 off_ids, u64 file_bytes
//   table    n_objects x {u64 first_row, u32 rows, f32 span, u32 id_offset, u32 id_len}     (24 B each)
//   descriptors  u8[total_rows][32]       objects concatenated in imgIdx order
//   points       f32[total_rows][3]
//   ids          the object_id strings, back to back (no terminators)
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "tod_internal.h"

namespace {

constexpr char kMagic[8] = {'T', 'O', 'D', 'B', '2', '0', '0', '\0'};
constexpr uint32_t kVersion = 1;

struct Header {
  char magic[8];
  uint32_t version;
  uint32_t n_objects;
  uint64_t total_rows;
  uint64_t off_table, off_desc, off_pts, off_ids, file_bytes;
};
static_assert(sizeof(Header) == 64, "header layout");

struct Entry {
  uint64_t first_row;
  uint32_t rows;
  float span;
  uint32_t id_offset, id_len;
};
static_assert(sizeof(Entry) == 24, "table entry layout");

inline uint64_t align64(uint64_t x) { return (x + 63) & ~uint64_t(63); }

// span of one object: DescriptorMatcher.cpp:106-121 (bounding-box diagonal, float arithmetic)
float object_span(const float *points, int32_t n) {
  float lo[3], hi[3];
  for (int d = 0; d < 3; ++d) {
    lo[d] = std::numeric_limits<float>::max();
    hi[d] = -std::numeric_limits<float>::max();
  }
  for (int32_t i = 0; i < n; ++i)
    for (int d = 0; d < 3; ++d) {
      lo[d] = std::min(lo[d], points[size_t(i) * 3 + d]);
      hi[d] = std::max(hi[d], points[size_t(i) * 3 + d]);
    }
  const float s = (hi[0] - lo[0]) * (hi[0] - lo[0]) + (hi[1] - lo[1]) * (hi[1] - lo[1]) + (hi[2] - lo[2]) * (hi[2] - lo[2]);
  return std::sqrt(s);
}

}  // namespace

struct tod_snapshot {
  void *base = nullptr;
  size_t bytes = 0;
  const Header *hdr = nullptr;
  const Entry *table = nullptr;
  std::vector<std::string> ids;
};

extern "C" {

int tod_snapshot_write(const char *path, int32_t n_objects, const char *const *object_ids,
                       const uint8_t *const *descriptors, const float *const *points, const int32_t *rows) {
  TOD_REQUIRE(path && n_objects >= 0 && (n_objects == 0 || (object_ids && descriptors && points && rows)),
              "null argument");
  uint64_t total = 0, id_bytes = 0;
  for (int32_t o = 0; o < n_objects; ++o) {
    TOD_REQUIRE(rows[o] >= 0 && object_ids[o] && (rows[o] == 0 || (descriptors[o] && points[o])), "bad object %d", o);
    total += uint64_t(rows[o]);
    id_bytes += std::strlen(object_ids[o]);
  }
  if (int64_t(total) > tod::kMaxDbRows)
    return tod::fail(TOD_ERR_LIMIT, "%llu descriptors exceed the %lld-row limit", (unsigned long long)total,
                     (long long)tod::kMaxDbRows);
  Header h{};
  std::memcpy(h.magic, kMagic, 8);
  h.version = kVersion;
  h.n_objects = uint32_t(n_objects);
  h.total_rows = total;
  h.off_table = 64;
  h.off_desc = align64(h.off_table + uint64_t(n_objects) * sizeof(Entry));
  h.off_pts = align64(h.off_desc + total * 32);
  h.off_ids = align64(h.off_pts + total * 12);
  h.file_bytes = h.off_ids + id_bytes;
  std::vector<unsigned char> buf(size_t(h.file_bytes), 0);
  std::memcpy(buf.data(), &h, sizeof(h));
  uint64_t row = 0, ido = 0;
  for (int32_t o = 0; o < n_objects; ++o) {
    Entry e{};
    e.first_row = row;
    e.rows = uint32_t(rows[o]);
    e.span = object_span(points[o], rows[o]);
    e.id_offset = uint32_t(ido);
    e.id_len = uint32_t(std::strlen(object_ids[o]));
    std::memcpy(buf.data() + h.off_table + size_t(o) * sizeof(Entry), &e, sizeof(e));
    if (rows[o]) {
      std::memcpy(buf.data() + h.off_desc + row * 32, descriptors[o], size_t(rows[o]) * 32);
      std::memcpy(buf.data() + h.off_pts + row * 12, points[o], size_t(rows[o]) * 12);
    }
    std::memcpy(buf.data() + h.off_ids + ido, object_ids[o], e.id_len);
    row += uint64_t(rows[o]);
    ido += e.id_len;
  }
  const std::string tmp = std::string(path) + ".tmp";
  FILE *f = std::fopen(tmp.c_str(), "wb");
  if (!f) return tod::fail(TOD_ERR_INVALID, "cannot create %s", tmp.c_str());
  const bool ok = std::fwrite(buf.data(), 1, buf.size(), f) == buf.size();
  const bool closed = std::fclose(f) == 0;
  if (!ok || !closed || std::rename(tmp.c_str(), path) != 0) {
    std::remove(tmp.c_str());
    return tod::fail(TOD_ERR_INVALID, "cannot write %s", path);
  }
  return TOD_OK;
}

int tod_snapshot_open(const char *path, tod_snapshot **out) {
  TOD_REQUIRE(path && out, "null argument");
  const int fd = ::open(path, O_RDONLY);
  if (fd < 0) return tod::fail(TOD_ERR_INVALID, "cannot open %s", path);
  struct stat st;
  if (fstat(fd, &st) != 0 || size_t(st.st_size) < sizeof(Header)) {
    ::close(fd);
    return tod::fail(TOD_ERR_PARSE, "%s is not a TOD DB snapshot (too short)", path);
  }
  void *base = mmap(nullptr, size_t(st.st_size), PROT_READ, MAP_PRIVATE, fd, 0);
  ::close(fd);
  if (base == MAP_FAILED) return tod::fail(TOD_ERR_INVALID, "mmap of %s failed", path);
  const Header *h = static_cast<const Header *>(base);
  const uint64_t n = h->n_objects, rows = h->total_rows, size = uint64_t(st.st_size);
  // Every bound is checked by subtraction against the file size (no u64 sum can wrap): the offsets are ordered,
  // aligned, inside the file, and each region is large enough for the counts the header claims.
  bool ok = std::memcmp(h->magic, kMagic, 8) == 0 && h->version == kVersion && h->file_bytes == size &&
            int64_t(rows) <= tod::kMaxDbRows && n <= size / sizeof(Entry);
  ok = ok && h->off_table >= sizeof(Header) && h->off_table <= size && h->off_table % 8 == 0 &&
       h->off_desc <= size && h->off_pts <= size && h->off_ids <= size && h->off_table <= h->off_desc &&
       h->off_desc <= h->off_pts && h->off_pts <= h->off_ids && h->off_desc % 4 == 0 && h->off_pts % 4 == 0;
  ok = ok && n <= (h->off_desc - h->off_table) / sizeof(Entry) && rows <= (h->off_pts - h->off_desc) / 32 &&
       rows <= (h->off_ids - h->off_pts) / 12;
  tod_snapshot *s = new tod_snapshot();
  s->base = base;
  s->bytes = size_t(st.st_size);
  s->hdr = h;
  if (ok) {
    s->table = reinterpret_cast<const Entry *>(static_cast<const char *>(base) + h->off_table);
    uint64_t row = 0;
    for (uint64_t o = 0; o < n && ok; ++o) {
      const Entry &e = s->table[o];
      ok = e.first_row == row && h->off_ids + uint64_t(e.id_offset) + e.id_len <= size;
      row += e.rows;
      if (ok) s->ids.emplace_back(static_cast<const char *>(base) + h->off_ids + e.id_offset, e.id_len);
    }
    ok = ok && row == rows;
  }
  if (!ok) {
    munmap(base, size_t(st.st_size));
    delete s;
    return tod::fail(TOD_ERR_PARSE, "%s is not a valid TOD DB snapshot (version %u expected)", path, kVersion);
  }
  *out = s;
  return TOD_OK;
}

void tod_snapshot_close(tod_snapshot *s) {
  if (!s) return;
  if (s->base) munmap(s->base, s->bytes);
  delete s;
}

int32_t tod_snapshot_num_objects(const tod_snapshot *s) { return s ? int32_t(s->hdr->n_objects) : 0; }
int64_t tod_snapshot_num_descriptors(const tod_snapshot *s) { return s ? int64_t(s->hdr->total_rows) : 0; }

int tod_snapshot_object(const tod_snapshot *s, int32_t index, const char **object_id, const uint8_t **descriptors,
                        const float **points, int32_t *rows, float *span) {
  TOD_REQUIRE(s, "null argument");
  TOD_REQUIRE(index >= 0 && uint32_t(index) < s->hdr->n_objects, "object index %d out of range", index);
  const Entry &e = s->table[index];
  const char *base = static_cast<const char *>(s->base);
  if (object_id) *object_id = s->ids[size_t(index)].c_str();
  if (descriptors) *descriptors = reinterpret_cast<const uint8_t *>(base + s->hdr->off_desc + e.first_row * 32);
  if (points) *points = reinterpret_cast<const float *>(base + s->hdr->off_pts + e.first_row * 12);
  if (rows) *rows = int32_t(e.rows);
  if (span) *span = e.span;
  return TOD_OK;
}

// parameter_callback from a snapshot: clear, add every object in file order (= imgIdx order).  The caller trains.
int tod_matcher_load_snapshot(tod_matcher *m, const char *path) {
  TOD_REQUIRE(m && path, "null argument");
  tod_snapshot *s = nullptr;
  if (int rc = tod_snapshot_open(path, &s)) return rc;
  int rc = tod_matcher_clear(m);
  for (int32_t o = 0; rc == TOD_OK && o < tod_snapshot_num_objects(s); ++o) {
    const char *id;
    const uint8_t *d;
    const float *p;
    int32_t rows;
    rc = tod_snapshot_object(s, o, &id, &d, &p, &rows, nullptr);
    if (rc == TOD_OK) rc = tod_matcher_add_object(m, id, d, p, rows);
  }
  tod_snapshot_close(s);
  return rc;
}

}  // extern "C"
