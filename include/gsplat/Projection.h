// Projection.h -- the reference's rasterizer_autograd.cpp includes "Projection.h"
// (src/training/rasterization/rasterizer_autograd.cpp:6) without using anything from it: the
// launcher it declares (gsplat/Projection.h:12) is internal to the reference's gsplat library.
// Kept as an empty include so that translation unit compiles unchanged against this backend.
#pragma once
#include "Ops.h"
