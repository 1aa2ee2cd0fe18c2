// refshim: MVE math/defines.h stand-in (see ../README.md)
#pragma once
#define MATH_NAMESPACE_BEGIN namespace math {
#define MATH_NAMESPACE_END }
#define MATH_PI 3.14159265358979323846264338327950288
#define MATH_DEG2RAD(x) ((x) * (MATH_PI / 180.0))
#define MATH_RAD2DEG(x) ((x) * (180.0 / MATH_PI))
#define MATH_POW2(x) ((x) * (x))
