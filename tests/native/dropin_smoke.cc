// Compiles the drop-in C++ classes (include/ORBextractor.h, include/ORBmatcher.h) the way the
// reference's Frame.cc would use them — against an OpenCV-API (here the test shim; a real build
// uses OpenCV) — and runs them on one image read from stdin-free binary files.
//   dropin_smoke <image.bin: int32 w, int32 h, bytes> <out.bin>
// Output: int32 n, n x orbx_keypoint, n x 32 descriptor bytes, int32 self-match count.
#include <cstdio>
#include <map>
#include <vector>

#include "ORBextractor.h"
#include "ORBmatcher.h"

struct MiniFrame {  // the members of ORB_SLAM2::Frame the matcher template reads
  std::vector<cv::KeyPoint> mvKeysUn;
  cv::Mat mDescriptors;
  static float mnMinX, mnMaxX, mnMinY, mnMaxY;
};
float MiniFrame::mnMinX = 0, MiniFrame::mnMaxX = 0, MiniFrame::mnMinY = 0, MiniFrame::mnMaxY = 0;

struct MiniMapPoint {
  bool isBad() const { return false; }
};
struct MiniKeyFrame {  // the members of ORB_SLAM2::KeyFrame / Frame that SearchByBoW_cam1 reads
  int N = 0;
  std::vector<cv::KeyPoint> mvKeysUn, mvKeys;
  cv::Mat mDescriptors;
  std::map<unsigned int, std::vector<unsigned int>> mFeatVec_cam1;  // DBoW2::FeatureVector
  std::vector<MiniMapPoint*> points;
  std::vector<MiniMapPoint*> GetMapPointMatches_cam1() const { return points; }
};

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  FILE* f = fopen(argv[1], "rb");
  int wh[2];
  if (!f || fread(wh, 4, 2, f) != 2) return 3;
  cv::Mat img(wh[1], wh[0], CV_8UC1);
  if (fread(img.data, 1, (size_t)wh[0] * wh[1], f) != (size_t)wh[0] * wh[1]) return 3;
  fclose(f);
  ORB_SLAM2::ORBextractor ex(1000, 1.2f, 8, 20, 7);
  MiniFrame F1, F2;
  ex(img, cv::Mat(), F1.mvKeysUn, F1.mDescriptors);
  ex(img, cv::Mat(), F2.mvKeysUn, F2.mDescriptors);
  MiniFrame::mnMaxX = (float)wh[0];
  MiniFrame::mnMaxY = (float)wh[1];
  ORB_SLAM2::ORBmatcherB200 matcher(0.9f, true);
  std::vector<cv::Point2f> prev;
  for (const cv::KeyPoint& k : F1.mvKeysUn) prev.push_back(k.pt);
  std::vector<int> m12;
  const int nm = matcher.SearchForInitialization(F1, F2, prev, m12, 100);
  const int d01 = matcher.DescriptorDistance(F1.mDescriptors.rowRange(0, 1), F1.mDescriptors.rowRange(1, 2));
  // SearchByBoW_cam1 of the frame against itself: vocabulary node = octave, every keypoint holds a point
  MiniKeyFrame KF;
  KF.N = (int)F1.mvKeysUn.size();
  KF.mvKeysUn = KF.mvKeys = F1.mvKeysUn;
  KF.mDescriptors = F1.mDescriptors;
  std::vector<MiniMapPoint> store(KF.N);
  for (int i = 0; i < KF.N; ++i) {
    KF.points.push_back(&store[i]);
    KF.mFeatVec_cam1[(unsigned)F1.mvKeysUn[i].octave].push_back((unsigned)i);
  }
  std::vector<MiniMapPoint*> bow_matches;
  ORB_SLAM2::ORBmatcherB200 bow_matcher(0.7f, true);
  const int nbow = bow_matcher.SearchByBoW_cam1(&KF, KF, bow_matches);
  int bow_self = 0;
  for (int i = 0; i < KF.N; ++i) bow_self += bow_matches[i] == &store[i];
  FILE* o = fopen(argv[2], "wb");
  const int n = (int)F1.mvKeysUn.size();
  fwrite(&n, 4, 1, o);
  for (int i = 0; i < n; ++i) {
    const cv::KeyPoint& k = F1.mvKeysUn[i];
    orbx_keypoint kk = {k.pt.x, k.pt.y, k.size, k.angle, k.response, k.octave};
    fwrite(&kk, sizeof(kk), 1, o);
  }
  for (int i = 0; i < n; ++i) fwrite(F1.mDescriptors.ptr(i), 1, 32, o);
  fwrite(&nm, 4, 1, o);
  fwrite(&d01, 4, 1, o);
  int self = 0;
  for (int i = 0; i < (int)m12.size(); ++i) self += m12[i] == i;
  fwrite(&self, 4, 1, o);
  const int pw = ex.mvImagePyramid[3].cols, ph = ex.mvImagePyramid[3].rows;
  fwrite(&pw, 4, 1, o);
  fwrite(&ph, 4, 1, o);
  fwrite(&nbow, 4, 1, o);
  fwrite(&bow_self, 4, 1, o);
  fclose(o);
  printf("dropin: %d keypoints, %d init matches (%d self), d01=%d, pyramid[3]=%dx%d, %d BoW matches (%d self)\n", n, nm,
         self, d01, pw, ph, nbow, bow_self);
  return 0;
}
