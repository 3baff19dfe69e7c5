// Stand-in for the reference's precompiled header (src/pch.h) when compiling
// its scene-prep sources in place from /root/reference: only what
// src/resources/meshlets.cpp and src/shared_cpu_gpu.h need.  TEST
// INFRASTRUCTURE (oracle/_ref build); written for this repo, not copied.
#pragma once
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <vector>

#include <glm/glm.hpp>
#include <meshoptimizer.h>
#define dbg(...) ((void)0)
