/* TEST INFRASTRUCTURE — see phys_impl.h.  Builds the f32 and f64 instances of the physics oracle. */
#include <math.h>
#include <stddef.h>

#define REAL float
#define SFX _f32
#define SQRT sqrtf
#define SIN sinf
#define COS cosf
#define FLOOR floorf
#define FABS fabsf
#include "phys_impl.h"
#undef REAL
#undef SFX
#undef SQRT
#undef SIN
#undef COS
#undef FLOOR
#undef FABS

#define REAL double
#define SFX _f64
#define SQRT sqrt
#define SIN sin
#define COS cos
#define FLOOR floor
#define FABS fabs
#include "phys_impl.h"
