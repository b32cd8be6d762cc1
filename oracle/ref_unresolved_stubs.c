/* ORACLE/_ref - TEST INFRASTRUCTURE ONLY.  Link-time placeholders for functions that the compiled reference translation units (oracle/ref_assoc_wrap.cpp) CALL but
 * that the reference DEFINES in files which are not compiled here (drawing, SIFT, triangulation, image line detection / matching - real OpenCV algorithms, all outside
 * the hot path).  Each placeholder carries the exact mangled name so the library loads, and aborts loudly if anything ever reaches it: no exported entry point does. */
#include <stdio.h>
#include <stdlib.h>
static void not_compiled(const char* what) { fprintf(stderr, "oracle/_ref: %s belongs to a reference file that is not part of this build\n", what); abort(); }

int _Z15CameraCenterPCDRKNSt7__cxx1112basic_stringIcSt11char_traitsIcESaIcEEERKSt6vectorIN5Eigen6MatrixIdLi3ELi1EEENS8_17aligned_allocatorISA_EEE(void) { return 0; }   /* a debug file writer (util/Visualization.cpp) called unconditionally by JointOptimize: the stand-in writes nothing */
void _Z16DrawLinesOnImageRKN2cv3MatERKSt6vectorINS_3VecIfLi4EEESaIS5_EERKS3_INS_6ScalarESaISA_EEibbRKS3_IiSaIiEE(void) { not_compiled("DrawLinesOnImage"); }
void _Z16TriangulateNViewRKSt6vectorIN5Eigen6MatrixIdLi3ELi3EEENS0_17aligned_allocatorIS2_EEERKS_INS1_IdLi3ELi1EEENS3_IS8_EEERKS_IN2cv7Point3_IfEESaISF_EE(void) { not_compiled("TriangulateNView"); }
int _Z19CameraPoseVisualizeRKNSt7__cxx1112basic_stringIcSt11char_traitsIcESaIcEEERKSt6vectorIN5Eigen6MatrixIdLi3ELi3EEENS8_17aligned_allocatorISA_EEERKS7_INS9_IdLi3ELi1EEENSB_ISG_EEEi(void) { return 0; }   /* a debug file writer (util/Visualization.cpp) called unconditionally by JointOptimize: the stand-in writes nothing */
void _Z19ExtractSIFTQuadtreeRKN2cv3MatERSt6vectorINS_8KeyPointESaIS4_EEiiS2_(void) { not_compiled("ExtractSIFTQuadtree"); }
void _Z20DrawLinePairsOnImageRKN2cv3MatERKSt6vectorI19CameraLidarLinePairSaIS4_EERKN5Eigen6MatrixIdLi4ELi4EEEib(void) { not_compiled("DrawLinePairsOnImage"); }
void _Z21ComputeSIFTDescriptorRKN2cv3MatERSt6vectorINS_8KeyPointESaIS4_EERS0_b(void) { not_compiled("ComputeSIFTDescriptor"); }
void _ZN12PanoramaLine4FuseEfb(void) { not_compiled("PanoramaLine::Fuse"); }
void _ZN12PanoramaLine6DetectERKN2cv3MatE(void) { not_compiled("PanoramaLine::Detect"); }
void _ZN12PanoramaLine6DetectEff(void) { not_compiled("PanoramaLine::Detect"); }
void _ZN12PanoramaLineC1ERKN2cv3MatEi(void) { not_compiled("PanoramaLine::PanoramaLine"); }
void _ZN19PanoramaLineMatcher14GenerateTracksEi(void) { not_compiled("PanoramaLineMatcher::GenerateTracks"); }
void _ZN19PanoramaLineMatcher15SetNeighborSizeEi(void) { not_compiled("PanoramaLineMatcher::SetNeighborSize"); }
void _ZN19PanoramaLineMatcher17SetMinTrackLengthEi(void) { not_compiled("PanoramaLineMatcher::SetMinTrackLength"); }
void _ZN19PanoramaLineMatcher19RemoveParallelLinesEv(void) { not_compiled("PanoramaLineMatcher::RemoveParallelLines"); }
void _ZN19PanoramaLineMatcherC1ERKSt6vectorI12PanoramaLineSaIS1_EERKS0_IN5Eigen6MatrixIdLi3ELi3EEENS6_17aligned_allocatorIS8_EEERKS0_INS7_IdLi3ELi1EEENS9_ISE_EEE(void) { not_compiled("PanoramaLineMatcher::PanoramaLineMatcher"); }
void _ZNK19PanoramaLineMatcher9GetTracksEv(void) { not_compiled("PanoramaLineMatcher::GetTracks"); }
