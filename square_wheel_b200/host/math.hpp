// Small vector / matrix library with cgmath 0.18 conventions (column-major Mat4, m * v = sum of columns), as used by the
// reference through src/common/math_types.rs.  Host-side only: everything computed here (matrices, equations, clipped
// vertices) is INPUT to the kernels, so it does not have to reproduce cgmath bit for bit.
#pragma once
#include <cmath>
#include <cstdint>

namespace sqwh
{

struct Vec2f
{
	float x = 0, y = 0;
};
struct Vec3f
{
	float x = 0, y = 0, z = 0;
	float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
	float& at(int i) { return i == 0 ? x : (i == 1 ? y : z); }
};
struct Vec4f
{
	float x = 0, y = 0, z = 0, w = 0;
};

inline Vec2f operator+(Vec2f a, Vec2f b) { return {a.x + b.x, a.y + b.y}; }
inline Vec2f operator-(Vec2f a, Vec2f b) { return {a.x - b.x, a.y - b.y}; }
inline Vec2f operator*(Vec2f a, float s) { return {a.x * s, a.y * s}; }
inline Vec2f operator/(Vec2f a, float s) { return {a.x / s, a.y / s}; }
inline float dot(Vec2f a, Vec2f b) { return a.x * b.x + a.y * b.y; }

inline Vec3f operator+(Vec3f a, Vec3f b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3f operator-(Vec3f a, Vec3f b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3f operator-(Vec3f a) { return {-a.x, -a.y, -a.z}; }
inline Vec3f operator*(Vec3f a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline Vec3f operator/(Vec3f a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline float dot(Vec3f a, Vec3f b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vec3f cross(Vec3f a, Vec3f b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline float magnitude(Vec3f a) { return std::sqrt(dot(a, a)); }
inline Vec3f normalize(Vec3f a) { return a / magnitude(a); }

inline Vec4f extend(Vec3f v, float w) { return {v.x, v.y, v.z, w}; }
inline Vec3f truncate(Vec4f v) { return {v.x, v.y, v.z}; }
inline Vec2f truncate(Vec3f v) { return {v.x, v.y}; }
inline Vec4f operator+(Vec4f a, Vec4f b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline Vec4f operator*(Vec4f a, float s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
inline float dot(Vec4f a, Vec4f b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

struct Plane // common/plane.rs: points with dot(vec, p) == dist
{
	Vec3f vec;
	float dist = 0;
};

struct Mat4f
{
	Vec4f x, y, z, w; // columns
	static Mat4f identity() { return {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}}; }
	static Mat4f from_cols(Vec4f a, Vec4f b, Vec4f c, Vec4f d) { return {a, b, c, d}; }
	static Mat4f from_translation(Vec3f v)
	{
		Mat4f m = identity();
		m.w = {v.x, v.y, v.z, 1};
		return m;
	}
	static Mat4f from_nonuniform_scale(float sx, float sy, float sz)
	{
		Mat4f m = identity();
		m.x.x = sx;
		m.y.y = sy;
		m.z.z = sz;
		return m;
	}
	Vec4f row(int i) const
	{
		switch (i)
		{
		case 0: return {x.x, y.x, z.x, w.x};
		case 1: return {x.y, y.y, z.y, w.y};
		case 2: return {x.z, y.z, z.z, w.z};
		default: return {x.w, y.w, z.w, w.w};
		}
	}
	Mat4f transpose() const { return {row(0), row(1), row(2), row(3)}; }
	bool invert(Mat4f& out) const;
};
inline Vec4f operator*(const Mat4f& m, Vec4f v) { return m.x * v.x + m.y * v.y + m.z * v.z + m.w * v.w; }
inline Mat4f operator*(const Mat4f& a, const Mat4f& b) { return {a * b.x, a * b.y, a * b.z, a * b.w}; }

inline bool Mat4f::invert(Mat4f& out) const
{
	// cofactor expansion in double for robustness; result rounded to f32
	const double m[16] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w, z.x, z.y, z.z, z.w, w.x, w.y, w.z, w.w};
	double inv[16];
	inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
	inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
	inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
	inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
	inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
	inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
	inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
	inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
	inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
	inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
	inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
	inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
	inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
	inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
	inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
	inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
	const double det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
	if (det == 0.0 || !std::isfinite(det))
		return false;
	const double id = 1.0 / det;
	float r[16];
	for (int i = 0; i < 16; ++i)
		r[i] = (float)(inv[i] * id);
	out.x = {r[0], r[1], r[2], r[3]};
	out.y = {r[4], r[5], r[6], r[7]};
	out.z = {r[8], r[9], r[10], r[11]};
	out.w = {r[12], r[13], r[14], r[15]};
	return true;
}

struct Quaternion // cgmath Quaternion {s, v}
{
	float s = 1;
	Vec3f v;
	static Quaternion from_angle_z(float a) { return {std::cos(a * 0.5f), {0, 0, std::sin(a * 0.5f)}}; }
	static Quaternion from_angle_y(float a) { return {std::cos(a * 0.5f), {0, std::sin(a * 0.5f), 0}}; }
	Quaternion conjugate() const { return {s, -v}; }
};
inline Quaternion operator*(Quaternion a, Quaternion b)
{
	return {a.s * b.s - dot(a.v, b.v), Vec3f{a.s * b.v.x + a.v.x * b.s + a.v.y * b.v.z - a.v.z * b.v.y, a.s * b.v.y + a.v.y * b.s + a.v.z * b.v.x - a.v.x * b.v.z,
										   a.s * b.v.z + a.v.z * b.s + a.v.x * b.v.y - a.v.y * b.v.x}};
}
inline Mat4f mat4_from_quaternion(Quaternion q) // cgmath impl From<Quaternion> for Matrix4
{
	const float x2 = q.v.x + q.v.x, y2 = q.v.y + q.v.y, z2 = q.v.z + q.v.z;
	const float xx2 = x2 * q.v.x, xy2 = x2 * q.v.y, xz2 = x2 * q.v.z;
	const float yy2 = y2 * q.v.y, yz2 = y2 * q.v.z, zz2 = z2 * q.v.z;
	const float sy2 = y2 * q.s, sz2 = z2 * q.s, sx2 = x2 * q.s;
	return {{1 - yy2 - zz2, xy2 + sz2, xz2 - sy2, 0}, {xy2 - sz2, 1 - xx2 - zz2, yz2 + sx2, 0}, {xz2 + sy2, yz2 - sx2, 1 - xx2 - yy2, 0}, {0, 0, 0, 1}};
}

// common/matrix.rs:4-11
struct CameraMatrices
{
	Vec3f position;
	Mat4f view_matrix;
	Mat4f planes_matrix;
};

// common/matrix.rs:82-110 complete_view_matrix
inline CameraMatrices complete_view_matrix(Vec3f position, const Mat4f& rotation_matrix, Vec2f fov_scale, float viewport_width, float viewport_height)
{
	const Mat4f translate = Mat4f::from_translation(-position);
	const Mat4f perspective = Mat4f::from_nonuniform_scale(fov_scale.x, fov_scale.y, 1.0f);
	const Mat4f resize_to_viewport = Mat4f::from_nonuniform_scale(viewport_width * 0.5f, viewport_height * 0.5f, 1.0f);
	Mat4f shift_to_viewport_center = Mat4f::identity();
	shift_to_viewport_center.z.x = viewport_width * 0.5f;
	shift_to_viewport_center.z.y = viewport_height * 0.5f;
	CameraMatrices r;
	r.position = position;
	r.view_matrix = shift_to_viewport_center * resize_to_viewport * perspective * rotation_matrix * translate;
	Mat4f inv;
	r.view_matrix.transpose().invert(inv);
	r.planes_matrix = inv;
	return r;
}

// common/matrix.rs:13-47, :66-80 build_view_matrix
inline CameraMatrices build_view_matrix(Vec3f position, float azimuth, float elevation, float fov, float viewport_width, float viewport_height)
{
	const Quaternion rotation = Quaternion::from_angle_z(azimuth + 3.14159265358979323846f * 0.5f) * Quaternion::from_angle_y(-elevation);
	const float inv_half_fov_tan = 1.0f / std::tan(fov * 0.5f);
	const float aspect = viewport_width / viewport_height;
	Mat4f basis_change = Mat4f::identity();
	basis_change.x = {0, 0, 1, 0};
	basis_change.y = {-1, 0, 0, 0};
	basis_change.z = {0, -1, 0, 0};
	const Mat4f rot = basis_change * mat4_from_quaternion(rotation.conjugate());
	return complete_view_matrix(position, rot, Vec2f{inv_half_fov_tan / aspect, inv_half_fov_tan}, viewport_width, viewport_height);
}

inline Vec3f view_matrix_transform_vertex(const Mat4f& view_matrix, Vec3f v) { return truncate(view_matrix * extend(v, 1.0f)); } // matrix.rs:125

// PCG32 — every random choice in the synthetic workloads comes from this generator (seed in the result JSON).
struct Pcg32
{
	uint64_t state, inc;
	explicit Pcg32(uint64_t seed, uint64_t seq = 0x5157)
	{
		state = 0;
		inc = (seq << 1) | 1;
		next();
		state += seed;
		next();
	}
	uint32_t next()
	{
		const uint64_t old = state;
		state = old * 6364136223846793005ull + inc;
		const uint32_t xorshifted = (uint32_t)(((old >> 18) ^ old) >> 27);
		const uint32_t rot = (uint32_t)(old >> 59);
		return (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));
	}
	float uniform() { return (float)(next() >> 8) * (1.0f / 16777216.0f); }
	float uniform(float a, float b) { return a + (b - a) * uniform(); }
	uint32_t below(uint32_t n) { return n ? next() % n : 0; }
};

} // namespace sqwh
