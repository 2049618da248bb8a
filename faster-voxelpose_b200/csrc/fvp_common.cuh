// fvp_common.cuh - shared host/device declarations of libfvp_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/fvp_b200.h"

#define FVP_MAX_VIEWS 8
#define FVP_MAX_PEOPLE 16

// ---- calibration block of one view; the fp32 values the reference obtains from
// unfold_camera_param (lib/utils/cameras.py:11-18)
struct FvpCam {
  float R[9];
  float T[3];
  float fx, fy, cx, cy;
  float k[3];
  float p[2];
  float pad[3];
};
struct FvpSeq {
  FvpCam cam[FVP_MAX_VIEWS];
  float A[6];      // resize_transform, 2x3 row major
  float pad[2];
};

// constants of the voxel -> heat-map-pixel chain (project_whole.py:49-60) that do not depend on the view
struct FvpProj {
  float ori_max;       // max(ori_w, ori_h): the clamp bound used for BOTH coordinates
  float hm_w, hm_h;    // float(w), float(h)
  float img_w, img_h;
  float wm1, hm1;      // w-1, h-1
  float r_img_w, r_img_h, r_wm1, r_hm1;   // correctly rounded reciprocals of the four constant divisors
  int W, H;            // heat-map size
  int WP, HP;          // zero-bordered size
  int PADX, PADY;
  int JP;              // channels padded to a multiple of 4
};

// one proposal slot, produced by the proposal kernel, consumed by K3 / P2PNet / pose head
struct FvpPerson {
  int valid;           // flag >= 0 (human_detection_net.py:63 / faster_voxelpose.py:45)
  int empty;           // some start >= end: cube stays zero (project_individual.py:125)
  int tl[3];           // centers_tl (project_individual.py:110)
  int lo[3], hi[3];    // active cube index range [lo,hi) = [start-tl, end-tl)
  float offset[3];     // project_individual.py:111
  int seq;             // calibration slot of the frame
  int pad;
};

#define FVP_CUDA_OK(expr)                                                          \
  do {                                                                             \
    cudaError_t e__ = (expr);                                                      \
    if (e__ != cudaSuccess) return fvp_fail_cuda(ctx, e__, #expr, __FILE__, __LINE__); \
  } while (0)

struct fvp_ctx;
int fvp_fail_cuda(fvp_ctx* ctx, cudaError_t e, const char* what, const char* file, int line);
int fvp_fail(fvp_ctx* ctx, int code, const char* fmt, ...);

static inline int fvp_round_up(int a, int b) { return (a + b - 1) / b * b; }
static inline int fvp_cdiv(int a, int b) { return (a + b - 1) / b; }
