// rt_compat.cuh — the standard headers the kernel headers need, or their stand-ins when the same
// headers are compiled at run time by NVRTC (registered problems, rt_registry.cu): NVRTC ships no
// <stdint.h> / <math.h>; the math functions are built in.
#pragma once
#ifdef __CUDACC_RTC__
#include "../../include/b200dae.h"   // int32_t, int64_t
typedef signed char int8_t;
typedef short int16_t;
typedef unsigned char uint8_t;
typedef unsigned short uint16_t;
typedef unsigned int uint32_t;
typedef unsigned long long uint64_t;
#ifndef INFINITY
#define INFINITY (__longlong_as_double(0x7ff0000000000000LL))
#endif
#else
#include <math.h>
#include <stdint.h>
#endif
