// TEST INFRASTRUCTURE ONLY (oracle/): gcc-clean stand-in for the reference's vecmath.h.
//
// Why it exists: /root/reference/NBodySimulation/vecmath.h:56-62,84-99,123-161 nests
// types that have constructors inside anonymous structs, an MSVC extension g++ rejects.
// This header is force-included (-include) ahead of the reference sources and claims the
// reference's include guard, so sph.h / demo4.h / demo4.cpp compile *unmodified* against
// these types.  Only layout and arithmetic ORDER are reproduced (each function notes the
// reference line whose floating-point evaluation order it keeps); the text is ours.
//
// Nothing under nbodysimulation_experiment_b200/ may include this file.
#ifndef VECMATH_H
#define VECMATH_H

#define _USE_MATH_DEFINES
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <algorithm>

static const float kDeg2Rad = (float)M_PI / 180.0f;

// ---- storage types -------------------------------------------------------------------
// No user-declared operator= anywhere: demo4.h:62-67 keeps Plane/Circle/... in an anonymous
// union and assigns Body values, which needs trivially assignable members.
union Vec2i {
	struct { int x, y; };
	int m[2];
	Vec2i() : x(0), y(0) {}
	Vec2i(int ax, int ay) : x(ax), y(ay) {}
};

union Vec2f {
	struct { float x, y; };
	struct { float w, h; };
	float m[2];
	Vec2f() : x(0.0f), y(0.0f) {}
	Vec2f(float s) : x(s), y(s) {}                  // implicit on purpose (vecmath.h:46-49)
	Vec2f(float ax, float ay) : x(ax), y(ay) {}
};

struct Mat2f {                                       // nothing compiled here touches Mat2f::m
	Vec2f col1, col2;
	Mat2f() : col1(1.0f, 0.0f), col2(0.0f, 1.0f) {}
};

union Vec3f {
	struct { float x, y, z; };
	struct { float u, v, w; };
	struct { float r, g, b; };
	float m[3];
	Vec3f() : x(0), y(0), z(0) {}
	Vec3f(float s) : x(s), y(s), z(s) {}
	Vec3f(float ax, float ay, float az) : x(ax), y(ay), z(az) {}
};

union Vec4f {
	struct { float x, y, z, w; };
	struct { float r, g, b, a; };
	float m[4];
	Vec4f() : x(0), y(0), z(0), w(1) {}             // alpha defaults to 1 (vecmath.h:163-166)
	Vec4f(float ax, float ay, float az, float aw) : x(ax), y(ay), z(az), w(aw) {}
};

struct Mat4f {                                       // nothing compiled here touches Mat4f::m
	Vec4f col1, col2, col3, col4;
	Mat4f() : col1(1, 0, 0, 0), col2(0, 1, 0, 0), col3(0, 0, 1, 0), col4(0, 0, 0, 1) {}
	static Mat4f TransformationFromVec2(const Vec2f &p) {
		Mat4f t; t.col4.x = p.x; t.col4.y = p.y; t.col4.z = 0.0f; return t;
	}
	static Mat4f ScaleFromVec2(const Vec2f &s) {
		Mat4f t; t.col1.x = s.x; t.col2.y = s.y; t.col3.z = 0.0f; return t;
	}
};

union Pixel {
	struct { uint8_t r, g, b, a; };
	uint8_t m[4];
};

// ---- Vec2f arithmetic (component-wise; one rounding per component, like vecmath.h:233-260)
inline Vec2f operator*(const Vec2f &v, float s) { return Vec2f(v.x * s, v.y * s); }
inline Vec2f operator-(const Vec2f &v) { return Vec2f(-v.x, -v.y); }
inline Vec2f operator+(const Vec2f &p, const Vec2f &q) { return Vec2f(p.x + q.x, p.y + q.y); }
inline Vec2f operator-(const Vec2f &p, const Vec2f &q) { return Vec2f(p.x - q.x, p.y - q.y); }
inline Vec2f &operator*=(Vec2f &v, float s) { v = v * s; return v; }
// vecmath.h:249-252 evaluates `b + a`; addition commutes bit-for-bit in IEEE, kept anyway.
inline Vec2f &operator+=(Vec2f &v, const Vec2f &d) { v = d + v; return v; }
inline Vec2f &operator-=(Vec2f &v, const Vec2f &d) { v = v - d; return v; }

inline float ScalarLerp(float a, float t, float b) { return (1.0f - t) * a + t * b; }   // :228
inline float Vec2Dot(const Vec2f &p, const Vec2f &q) { return p.x * q.x + p.y * q.y; }  // :262
inline float Vec2Length(const Vec2f &v) { return sqrtf(v.x * v.x + v.y * v.y); }        // :267

// vecmath.h:272-280 — zero vector normalises with length 1 (stays zero); multiply by the
// reciprocal, do not divide.
inline Vec2f Vec2Normalize(const Vec2f &v) {
	float len = Vec2Length(v);
	if (len == 0) len = 1;
	float rcp = 1.0f / len;
	return v * rcp;
}

inline Vec2f Vec2Hadamard(const Vec2f &p, const Vec2f &q) { return Vec2f(p.x * q.x, p.y * q.y); }

inline Vec2f Vec2MultMat2(const Mat2f &A, const Vec2f &v) {                             // :287
	return Vec2f(A.col1.x * v.x + A.col2.x * v.y, A.col1.y * v.x + A.col2.y * v.y);
}

// vecmath.h:292-296 — NOT a distance: (dx*dy)^2.  The polygon vertex regions depend on it
// (sph.h:637,648), so the quirk is part of the behaviour under test.
inline float Vec2DistanceSquared(const Vec2f &p, const Vec2f &q) {
	float prod = (q.x - p.x) * (q.y - p.y);
	return prod * prod;
}

inline Vec2f Vec2Cross(const Vec2f &v, float s) { return Vec2f(s * v.y, -s * v.x); }    // right perp
inline Vec2f Vec2Cross(float s, const Vec2f &v) { return Vec2f(-s * v.y, s * v.x); }    // left perp
inline float Vec2Cross(const Vec2f &p, const Vec2f &q) { return p.x * q.y - p.y * q.x; }
inline float Vec2AxisToAngle(const Vec2f &axis) { return atan2f(axis.y, axis.x); }

// vecmath.h:317-322 — libc rand() stream; the particle jitter of every scenario hangs on it.
inline Vec2f Vec2RandomDirection() {
	float unit = rand() / (float)RAND_MAX;
	float ang = unit * ((float)M_PI * 2.0f);
	return Vec2f(cosf(ang), sinf(ang));
}

inline Vec2f Vec2Lerp(const Vec2f &p, float t, const Vec2f &q) {
	return Vec2f(ScalarLerp(p.x, t, q.x), ScalarLerp(p.y, t, q.y));
}

inline Vec3f operator*(float s, const Vec3f &v) { return Vec3f(s * v.x, s * v.y, s * v.z); }
inline Vec3f operator*(const Vec3f &v, float s) { return s * v; }
inline Vec3f &operator*=(Vec3f &v, float s) { v = s * v; return v; }

// ---- Mat2f ---------------------------------------------------------------------------
inline Mat2f Mat2Identity() { return Mat2f(); }
inline Mat2f Mat2FromAngle(float angle) {                                               // :358
	float s = sinf(angle), c = cosf(angle);
	Mat2f r; r.col1 = Vec2f(c, s); r.col2 = Vec2f(-s, c); return r;
}
inline Mat2f Mat2FromAxis(const Vec2f &axis) {                                          // :367
	Mat2f r; r.col1 = axis; r.col2 = Vec2Cross(1.0f, axis); return r;
}
inline Mat2f Mat2Transpose(const Mat2f &a) {
	Mat2f r; r.col1 = Vec2f(a.col1.x, a.col2.x); r.col2 = Vec2f(a.col1.y, a.col2.y); return r;
}
inline Mat2f Mat2Mult(const Mat2f &a, const Mat2f &b) {
	Mat2f r; r.col1 = Vec2MultMat2(a, b.col1); r.col2 = Vec2MultMat2(a, b.col2); return r;
}
inline float Mat2ToAngle(const Mat2f &a) { return Vec2AxisToAngle(a.col1); }
inline Mat2f Mat2MultTranspose(const Mat2f &a, const Mat2f &b) {
	Mat2f r;
	r.col1 = Vec2f(Vec2Dot(a.col1, b.col1), Vec2Dot(a.col2, b.col1));
	r.col2 = Vec2f(Vec2Dot(a.col1, b.col2), Vec2Dot(a.col2, b.col2));
	return r;
}

// ---- colours (render.h / demo4.cpp only name these) -----------------------------------
static const Vec4f ColorWhite = Vec4f(1.0f, 1.0f, 1.0f, 1.0f);
static const Vec4f ColorRed = Vec4f(1.0f, 0.0f, 0.0f, 1.0f);
static const Vec4f ColorGreen = Vec4f(0.0f, 1.0f, 0.0f, 1.0f);
static const Vec4f ColorBlue = Vec4f(0.0f, 0.0f, 1.0f, 1.0f);
static const Vec4f ColorLightGray = Vec4f(0.3f, 0.3f, 0.3f, 1.0f);
static const Vec4f ColorDarkGray = Vec4f(0.2f, 0.2f, 0.2f, 1.0f);

static const float INV255 = 1.0f / 255.0f;
inline Pixel RGBA32ToPixel(uint32_t v) {
	Pixel p; p.r = v & 0xFF; p.g = (v >> 8) & 0xFF; p.b = (v >> 16) & 0xFF; p.a = (v >> 24) & 0xFF; return p;
}
inline uint32_t RGBA32(uint8_t r, uint8_t g, uint8_t b, uint8_t a) {
	return ((uint32_t)a << 24) | ((uint32_t)b << 16) | ((uint32_t)g << 8) | (uint32_t)r;
}
inline Vec4f PixelToLinear(const Pixel &p) { return Vec4f(p.r * INV255, p.g * INV255, p.b * INV255, p.a * INV255); }
inline Vec4f RGBA32ToLinear(uint32_t v) { return PixelToLinear(RGBA32ToPixel(v)); }
inline Vec4f AlphaToLinear(uint8_t alpha) { return Vec4f(1, 1, 1, alpha * INV255); }
inline uint32_t LinearToRGBA32(const Vec4f &c) {
	return RGBA32((uint8_t)(c.x * 255.0f + 0.5f), (uint8_t)(c.y * 255.0f + 0.5f),
	              (uint8_t)(c.z * 255.0f + 0.5f), (uint8_t)(c.w * 255.0f + 0.5f));
}

#endif // VECMATH_H
