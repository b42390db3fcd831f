// Stand-in for deps:dso/src/util/settings.h: the globals the hot path reads (values: settings.cpp:127, 138; mode 1 of src/main.cpp:117-122)
#pragma once
#define PYR_LEVELS 6
namespace dso {
extern int pyrLevelsUsed;
extern int wG[PYR_LEVELS], hG[PYR_LEVELS];
extern float setting_huberTH;
extern float setting_coarseCutoffTH;
extern float setting_affineOptModeA;
extern float setting_affineOptModeB;
extern bool setting_debugout_runquiet;
extern int setting_gammaWeightsPixelSelect;  // 1 = weight absSquaredGrad by the response gradient (settings.cpp)
extern float freeDebugParam3;  // only read by colour-map helpers of util/globalFuncs.h
}  // namespace dso
