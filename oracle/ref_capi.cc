// TEST INFRASTRUCTURE — CPU oracle, not product code.
//
// C API (oracle/orb_oracle.h, extractor half) over the reference's own ORB_SLAM2::ORBextractor,
// whose source is compiled verbatim from /root/reference/src/ORBextractor.cc (never copied into
// this repo) against oracle/shim.  Also provides the monotonic bump allocator that pins the
// pointer-address tie-break of ORBextractor.cc:685: node addresses grow in
// creation order, so equal-size nodes are split "later-created first", deterministically.
#include <sys/mman.h>

#include <cstdio>
#include <cstdlib>
#include <new>

#include "ORBextractor.h"  // from /root/reference/include
#include "orb_oracle.h"

namespace {
struct Arena {
  char* base = nullptr;
  size_t cap = 0, off = 0;
};
thread_local Arena g_arena;
const size_t kArenaBytes = (size_t)8 << 30;  // virtual reservation; pages are touched lazily

inline void* arena_alloc(size_t n) {
  Arena& a = g_arena;
  if (!a.base) {
    void* p = mmap(nullptr, kArenaBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (p == MAP_FAILED) { std::fprintf(stderr, "orb_ref: arena mmap failed\n"); std::abort(); }
    a.base = (char*)p;
    a.cap = kArenaBytes;
  }
  n = (n + 15) & ~(size_t)15;
  if (a.off + n > a.cap) { std::fprintf(stderr, "orb_ref: arena exhausted\n"); std::abort(); }
  void* r = a.base + a.off;
  a.off += n;
  return r;
}
}  // namespace

// Kept private to this library by -Bsymbolic + the ref.map version script.
#define HID
HID void* operator new(size_t n) { return arena_alloc(n); }
HID void* operator new[](size_t n) { return arena_alloc(n); }
HID void operator delete(void*) noexcept {}
HID void operator delete[](void*) noexcept {}
HID void operator delete(void*, size_t) noexcept {}
HID void operator delete[](void*, size_t) noexcept {}

struct oo_extractor {
  int nfeatures, nlevels, ini_th, min_th;
  float scale_factor;
  ORB_SLAM2::ORBextractor* ex;  // lives in the arena for the duration of one call
  std::vector<cv::Mat>* pyr;
};

extern "C" {

oo_extractor* oo_create(int nf, float sf, int nl, int ini, int mn) {
  oo_extractor* e = (oo_extractor*)std::malloc(sizeof(oo_extractor));
  e->nfeatures = nf; e->nlevels = nl; e->ini_th = ini; e->min_th = mn; e->scale_factor = sf;
  e->ex = nullptr; e->pyr = nullptr;
  return e;
}
void oo_destroy(oo_extractor* e) { std::free(e); }

int oo_extract(oo_extractor* e, const uint8_t* img, int rows, int cols, size_t stride,
               oo_keypoint* kps, uint8_t* desc, int cap, int* level_counts) {
  g_arena.off = 0;  // everything from the previous call is dead: restart the monotonic heap
  e->ex = new ORB_SLAM2::ORBextractor(e->nfeatures, e->scale_factor, e->nlevels, e->ini_th, e->min_th);
  cv::Mat image(rows, cols, CV_8UC1, (void*)img, stride);
  std::vector<cv::KeyPoint> keys;
  cv::Mat descriptors;
  (*e->ex)(image, cv::Mat(), keys, descriptors);
  e->pyr = &e->ex->mvImagePyramid;
  const int n = (int)keys.size();
  if (level_counts) for (int l = 0; l < e->nlevels; ++l) level_counts[l] = 0;
  if (n > cap) return -1;
  for (int i = 0; i < n; ++i) {
    kps[i].x = keys[i].pt.x; kps[i].y = keys[i].pt.y; kps[i].size = keys[i].size;
    kps[i].angle = keys[i].angle; kps[i].response = keys[i].response; kps[i].octave = keys[i].octave;
    if (level_counts) level_counts[keys[i].octave]++;
    std::memcpy(desc + (size_t)i * 32, descriptors.ptr(i), 32);
  }
  return n;
}

int oo_pyramid_level(oo_extractor* e, int level, const uint8_t** data, int* w, int* h, size_t* step) {
  if (!e->pyr || level < 0 || level >= e->nlevels) return -1;
  const cv::Mat& m = (*e->pyr)[level];
  *data = m.data; *w = m.cols; *h = m.rows; *step = m.step;
  return 0;
}

void oo_scale_tables(oo_extractor* e, float* s, float* is, float* s2, float* is2) {
  g_arena.off = 0;
  ORB_SLAM2::ORBextractor ex(e->nfeatures, e->scale_factor, e->nlevels, e->ini_th, e->min_th);
  e->pyr = nullptr;
  std::vector<float> a = ex.GetScaleFactors(), b = ex.GetInverseScaleFactors(),
                     c = ex.GetScaleSigmaSquares(), d = ex.GetInverseScaleSigmaSquares();
  for (int i = 0; i < e->nlevels; ++i) {
    if (s) s[i] = a[i];
    if (is) is[i] = b[i];
    if (s2) s2[i] = c[i];
    if (is2) is2[i] = d[i];
  }
}

}  // extern "C"
