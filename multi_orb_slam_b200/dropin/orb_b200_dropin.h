/* orb_b200_dropin.h — optional controls of the drop-in translation units (ORBextractor_b200.cc,
 * ORBmatcher_b200.cc).  The drop-ins define the members of the reference's OWN classes
 * (include/ORBextractor.h:45-112, include/ORBmatcher.h:37-137 — both headers stay untouched), so nothing here is
 * needed to build; these functions only expose what the reference's class layout has no room for. */
#ifndef ORB_B200_DROPIN_H
#define ORB_B200_DROPIN_H

namespace ORB_SLAM2 {
class ORBextractor;
namespace b200 {

/* CUDA device the drop-ins create their handles on (-1 = the current device; default).  Call before the first
 * extraction / match of the process. */
void SetDevice(int device);

/* ORBextractor::mvImagePyramid (public member, include/ORBextractor.h:92) is NOT mirrored to host memory by
 * default: no live code of the reference reads it (its only reader, Frame::ComputeStereoMatches, is commented
 * out, src/Frame.cc:782-956) and the download costs as much as the extraction.  mirror = true makes operator()
 * fill it after every call exactly as the reference lays it out: level l is a ROI inside a
 * (w + 2*19) x (h + 2*19) parent whose border is BORDER_REFLECT_101 (src/ORBextractor.cc:1109-1134). */
void SetPyramidMirror(ORBextractor* extractor, bool mirror);

/* The reference's `~ORBextractor(){}` is inline in its header, so the drop-in cannot hook destruction: the GPU
 * handle of an extractor lives until Release(extractor) or process exit. */
void Release(ORBextractor* extractor);

}  // namespace b200
}  // namespace ORB_SLAM2
#endif
