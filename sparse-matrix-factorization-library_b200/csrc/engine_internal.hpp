// What the other device-side translation units of the library (zengine.cu) may see of a plan.  Not part of the C ABI.
#pragma once
#include <cuda_runtime.h>

#include "plan.hpp"

struct sfb200_plan;

namespace sfb200 {

struct PlanInternals {
    const HostPlan *H;
    double *d_X;            // image of the padded factor
    double *d_lx, *d_ux;    // matrix values of the plan (destination of every upload; d_ux null unless the plan is split)
    cudaStream_t stream;
    int device, sms;
};

PlanInternals plan_internals(sfb200_plan *p);      // engine.cu

}  // namespace sfb200
