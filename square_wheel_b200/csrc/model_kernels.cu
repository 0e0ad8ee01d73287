// Triangle-model preparation on sm_100a.
//
// Replaces the body of PartialRenderer::build_dynamic_models_buffers (src/square_wheel/partial_renderer.rs:1122-1190) — per visible
// mesh: animate_and_transform_triangle_mesh_vertices (triangle_models_rendering.rs:7-203: vertex / skeleton animation, transform
// into camera space, texture-coordinate scale, per-vertex light from the light cube and two directional terms, :701-720),
// reject_triangle_model_back_faces (:327-345) and sort_model_triangles (:347-364).  What stays on the host is per-MODEL scalar work:
// matrices (cgmath), the light cube of the entity (light grid fetch + dynamic lights), the texture mip.
//
//   model_bones_kernel      one CTA per skeleton-animated instance: the bone chain `matrices[b] = matrices[parent] * lerp(frame
//                           bones)` (:113-140) walked in hierarchy order by one warp, 16 lanes = the 16 elements of a product;
//                           then one thread per bone for the weight-scaled view and normal matrices (:142-152).
//   model_vertices_kernel   one thread per vertex (grid.y = instance).
//   model_triangles_kernel  one CTA per instance: back-face flags, order-preserving compaction (ballot + block scan), then a
//                           bitonic sort of 64-bit keys (descending max z | mesh order), i.e. the stable order documented for
//                           the restatement: the reference's sort_unstable_by leaves ties implementation-defined.
//
// Every float expression is written in the operand order of square_wheel_b200/host/models.cpp / math.hpp and this file is
// compiled with -fmad=false, so vertices and triangle order are bit-identical to that mirror (tests/test_models_gpu.py).
#include "sqw_device.cuh"
#include "sqw_internal.h"

namespace sqw
{
namespace
{

struct V3
{
	float x, y, z;
};
struct V4
{
	float x, y, z, w;
};
struct M3 // columns
{
	V3 x, y, z;
};
struct M4 // columns
{
	V4 x, y, z, w;
};
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ V3 operator/(V3 a, float s) { return {__fdiv_rn(a.x, s), __fdiv_rn(a.y, s), __fdiv_rn(a.z, s)}; }
__device__ __forceinline__ float dot3(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross3(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ V4 operator+(V4 a, V4 b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
__device__ __forceinline__ V4 operator*(V4 a, float s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
__device__ __forceinline__ V4 operator*(const M4& m, V4 v) { return m.x * v.x + m.y * v.y + m.z * v.z + m.w * v.w; } // math.hpp: sum of columns
__device__ __forceinline__ M4 operator*(const M4& a, const M4& b) { return {a * b.x, a * b.y, a * b.z, a * b.w}; }
__device__ __forceinline__ M4 operator*(const M4& m, float s) { return {m.x * s, m.y * s, m.z * s, m.w * s}; }
__device__ __forceinline__ M4 operator+(const M4& a, const M4& b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
__device__ __forceinline__ V3 operator*(const M3& m, V3 v) { return m.x * v.x + m.y * v.y + m.z * v.z; }
__device__ __forceinline__ M3 operator*(const M3& m, float s) { return {m.x * s, m.y * s, m.z * s}; }
__device__ __forceinline__ M3 operator*(const M3& a, const M3& b) { return {a * b.x, a * b.y, a * b.z}; }
__device__ __forceinline__ M3 operator+(const M3& a, const M3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ M3 transpose3(const M3& m) { return {{m.x.x, m.y.x, m.z.x}, {m.x.y, m.y.y, m.z.y}, {m.x.z, m.y.z, m.z.z}}; }

// get_normals_matrix (triangle_models_rendering.rs:722-731): inverse of the transposed axis matrix (cgmath Matrix3::invert), identity
// when it is singular.
__device__ M3 normals_matrix_of(const M4& model_matrix)
{
	const M3 axis = {{model_matrix.x.x, model_matrix.x.y, model_matrix.x.z}, {model_matrix.y.x, model_matrix.y.y, model_matrix.y.z}, {model_matrix.z.x, model_matrix.z.y, model_matrix.z.z}};
	const M3 t = transpose3(axis);
	const V3 x = t.x, y = t.y, z = t.z;
	const float det = x.x * (y.y * z.z - z.y * y.z) - y.x * (x.y * z.z - z.y * x.z) + z.x * (x.y * y.z - y.y * x.z);
	if (det == 0.0f)
		return {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
	const M3 m = {cross3(y, z) / det, cross3(z, x) / det, cross3(x, y) / det};
	return transpose3(m);
}

__device__ __forceinline__ M4 load_m4(const float* p)
{
	return {{p[0], p[1], p[2], p[3]}, {p[4], p[5], p[6], p[7]}, {p[8], p[9], p[10], p[11]}, {p[12], p[13], p[14], p[15]}};
}

// light.rs:296-348 + triangle_models_rendering.rs:701-720
__device__ void vertex_light(const sqw_model_instance& in, V3 normal_transformed, float out[3])
{
	const V3 nn = normal_transformed * inv_sqrt_fast(fmaxf(dot3(normal_transformed, normal_transformed), 0.00000001f));
	out[0] = out[1] = out[2] = 0.0f;
	const float comp[3] = {nn.x, nn.y, nn.z};
#pragma unroll
	for (int axis = 0; axis < 3; ++axis)
	{
		if (comp[axis] <= 0.0f)
		{
#pragma unroll
			for (int i = 0; i < 3; ++i)
				out[i] += in.light_cube[axis * 2][i] * (-comp[axis]);
		}
		else
		{
#pragma unroll
			for (int i = 0; i < 3; ++i)
				out[i] += in.light_cube[axis * 2 + 1][i] * comp[axis];
		}
	}
#pragma unroll
	for (int d = 0; d < 2; ++d)
	{
		const V3 vs = {in.light_vector_scaled[d][0], in.light_vector_scaled[d][1], in.light_vector_scaled[d][2]};
		const float light_dir_dot = fmaxf(dot3(nn, vs), 0.0f);
#pragma unroll
		for (int i = 0; i < 3; ++i)
			out[i] += in.light_color[d][i] * light_dir_dot;
	}
}

__device__ __forceinline__ void store_vertex(float* dst, const M4& mv, V3 position, V3 normal_transformed, float tc0, float tc1, const sqw_model_instance& in)
{
	const V4 p = mv * V4{position.x, position.y, position.z, 1.0f}; // view_matrix_transform_vertex (matrix.rs:125)
	float light[3];
	vertex_light(in, normal_transformed, light);
	float4* o = reinterpret_cast<float4*>(dst);
	o[0] = make_float4(p.x, p.y, p.z, tc0 * in.tc_scale[0] + in.tc_shift[0]);
	o[1] = make_float4(tc1 * in.tc_scale[1] + in.tc_shift[1], light[0], light[1], light[2]);
}

constexpr int BONE_FLOATS = 16 + 9; // per bone: weight-scaled view matrix, weight-scaled normals matrix

__global__ void __launch_bounds__(256) model_bones_kernel(const sqw_model_instance* __restrict__ instances, const ModelMeshDev* __restrict__ meshes, float* __restrict__ bone_scratch)
{
	const sqw_model_instance& in = instances[blockIdx.x];
	const ModelMeshDev mesh = meshes[in.mesh];
	if (mesh.kind != SQW_MESH_SKELETON_ANIMATED || mesh.num_bones == 0 || mesh.frame_bones == nullptr)
		return;
	extern __shared__ float s_chain[]; // num_bones x 16: matrices[] of :113-140
	const uint32_t total_bones = mesh.num_bones, num_bones = min(total_bones, 255u);
	const float lerp0 = fminf(fmaxf(in.lerp, 0.0f), 1.0f), lerp1 = 1.0f - lerp0;
	const bool perform_lerp = lerp0 > 0.01f && lerp0 < 0.99f;
	const float* fb0 = mesh.frame_bones + (size_t)in.frames[0] * total_bones * 16;
	const float* fb1 = mesh.frame_bones + (size_t)in.frames[1] * total_bones * 16;
	if (threadIdx.x < 32)
	{
		// element e = column c, row r of the product: ((P.x[r] * m.c.x + P.y[r] * m.c.y) + P.z[r] * m.c.z) + P.w[r] * m.c.w
		const int e = threadIdx.x & 15, c = e >> 2, r = e & 3;
		for (uint32_t b = 0; b < num_bones; ++b)
		{
			float col[4];
#pragma unroll
			for (int k = 0; k < 4; ++k)
			{
				const float a0 = fb0[b * 16 + c * 4 + k], a1 = fb1[b * 16 + c * 4 + k];
				col[k] = perform_lerp ? a0 * lerp0 + a1 * lerp1 : (lerp0 > lerp1 ? a0 : a1);
			}
			const uint32_t parent = mesh.bone_parents[b];
			float v;
			if (parent < total_bones && parent < b)
			{
				const float* P = s_chain + parent * 16;
				v = P[0 * 4 + r] * col[0] + P[1 * 4 + r] * col[1] + P[2 * 4 + r] * col[2] + P[3 * 4 + r] * col[3];
			}
			else
				v = col[r];
			__syncwarp();
			if (threadIdx.x < 16)
				s_chain[b * 16 + e] = v;
			__syncwarp();
		}
	}
	__syncthreads();
	const M4 model_matrix = load_m4(in.model_matrix), model_view = load_m4(in.model_view_matrix);
	const float weight_scale = 1.0f / 255.0f;
	const M3 normals_matrix_weight_scaled = normals_matrix_of(model_matrix) * weight_scale;
	const M4 model_view_weight_scaled = model_view * weight_scale;
	for (uint32_t b = threadIdx.x; b < num_bones; b += blockDim.x)
	{
		const M4 m = load_m4(s_chain + b * 16);
		const M3 nm = normals_matrix_weight_scaled * normals_matrix_of(m);
		const M4 vm = model_view_weight_scaled * m;
		float* o = bone_scratch + ((size_t)in.bone_offset + b) * BONE_FLOATS;
		const float* vp = &vm.x.x;
#pragma unroll
		for (int k = 0; k < 16; ++k)
			o[k] = vp[k];
		const float* np = &nm.x.x;
#pragma unroll
		for (int k = 0; k < 9; ++k)
			o[16 + k] = np[k];
	}
}

__global__ void __launch_bounds__(128) model_vertices_kernel(
	const sqw_model_instance* __restrict__ instances, const ModelMeshDev* __restrict__ meshes, const float* __restrict__ bone_scratch, float* __restrict__ out_vertices)
{
	const sqw_model_instance& in = instances[blockIdx.y];
	const ModelMeshDev mesh = meshes[in.mesh];
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= mesh.num_vertices)
		return;
	float* dst = out_vertices + ((size_t)in.vertex_offset + i) * 8;
	const M4 model_view = load_m4(in.model_view_matrix);
	const float lerp0 = fminf(fmaxf(in.lerp, 0.0f), 1.0f), lerp1 = 1.0f - lerp0;
	const bool perform_lerp = lerp0 > 0.01f && lerp0 < 0.99f;
	if (mesh.kind == SQW_MESH_NON_ANIMATED)
	{
		const float* v = static_cast<const float*>(mesh.vertices) + (size_t)i * 8; // position, normal, tex_coord
		const M3 nm = normals_matrix_of(load_m4(in.model_matrix));
		store_vertex(dst, model_view, V3{v[0], v[1], v[2]}, nm * V3{v[3], v[4], v[5]}, v[6], v[7], in);
	}
	else if (mesh.kind == SQW_MESH_VERTEX_ANIMATED)
	{
		const float* f0 = static_cast<const float*>(mesh.vertices) + ((size_t)in.frames[0] * mesh.num_vertices + i) * 6; // position, normal
		const float* f1 = static_cast<const float*>(mesh.vertices) + ((size_t)in.frames[1] * mesh.num_vertices + i) * 6;
		const M3 nm = normals_matrix_of(load_m4(in.model_matrix));
		V3 position, normal;
		if (perform_lerp)
		{
			position = V3{f0[0], f0[1], f0[2]} * lerp0 + V3{f1[0], f1[1], f1[2]} * lerp1;
			normal = V3{f0[3], f0[4], f0[5]} * lerp0 + V3{f1[3], f1[4], f1[5]} * lerp1;
		}
		else
		{
			const float* f = lerp0 > lerp1 ? f0 : f1;
			position = V3{f[0], f[1], f[2]};
			normal = V3{f[3], f[4], f[5]};
		}
		store_vertex(dst, model_view, position, nm * normal, mesh.tex_coords[(size_t)i * 2], mesh.tex_coords[(size_t)i * 2 + 1], in);
	}
	else
	{
		// SkeletonAnimatedVertex: tex_coord[2], position, normal (f32), bone_index[4], weight[4] (u8): 40 bytes
		const char* vb = static_cast<const char*>(mesh.vertices) + (size_t)i * 40;
		const float* vf = reinterpret_cast<const float*>(vb);
		const uint8_t* bi = reinterpret_cast<const uint8_t*>(vb + 32);
		const uint8_t* wt = bi + 4;
		const V3 position = {vf[2], vf[3], vf[4]}, normal = {vf[5], vf[6], vf[7]};
		if (mesh.num_bones == 0 || mesh.frame_bones == nullptr)
		{
			const M3 nm = normals_matrix_of(load_m4(in.model_matrix));
			store_vertex(dst, model_view, position, nm * normal, vf[0], vf[1], in);
			return;
		}
		const uint32_t num_bones = min(mesh.num_bones, 255u);
		const float* bones = bone_scratch + (size_t)in.bone_offset * BONE_FLOATS;
		auto bone_m4 = [&](uint32_t b) { return load_m4(bones + (size_t)b * BONE_FLOATS); };
		auto bone_m3 = [&](uint32_t b) {
			const float* p = bones + (size_t)b * BONE_FLOATS + 16;
			return M3{{p[0], p[1], p[2]}, {p[3], p[4], p[5]}, {p[6], p[7], p[8]}};
		};
		const uint32_t i0 = min((uint32_t)bi[0], num_bones - 1);
		const float w0 = (float)wt[0];
		M4 mat = bone_m4(i0) * w0;
		M3 normal_mat = bone_m3(i0) * w0;
		for (int k = 1; k < 4; ++k)
			if (wt[k] > 0)
			{
				const uint32_t ii = min((uint32_t)bi[k], num_bones - 1);
				const float wi = (float)wt[k];
				mat = mat + bone_m4(ii) * wi;
				normal_mat = normal_mat + bone_m3(ii) * wi;
			}
		store_vertex(dst, mat, position, normal_mat * normal, vf[0], vf[1], in);
	}
}

__device__ __forceinline__ uint32_t descending_key(float z)
{
	if (z == 0.0f)
		z = 0.0f; // -0 and +0 compare equal in the reference's comparator
	const uint32_t b = __float_as_uint(z);
	const uint32_t ascending = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
	return ~ascending;
}

__global__ void __launch_bounds__(1024) model_triangles_kernel(
	const sqw_model_instance* __restrict__ instances, const ModelMeshDev* __restrict__ meshes, const float* __restrict__ vertices, unsigned long long* __restrict__ sort_scratch,
	uint16_t* __restrict__ out_triangles, uint32_t* __restrict__ out_counts)
{
	const sqw_model_instance& in = instances[blockIdx.x];
	const ModelMeshDev mesh = meshes[in.mesh];
	const float* v = vertices + (size_t)in.vertex_offset * 8;
	unsigned long long* keys = sort_scratch + in.sort_offset;
	__shared__ uint32_t s_warp[32];
	__shared__ uint32_t s_base;
	if (threadIdx.x == 0)
		s_base = 0;
	__syncthreads();
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	// reject_triangle_model_back_faces (:327-345) + the sort key of sort_model_triangles (:347-364), compacted in mesh order
	for (uint32_t t0 = 0; t0 < mesh.num_triangles; t0 += blockDim.x)
	{
		const uint32_t t = t0 + threadIdx.x;
		bool keep = false;
		float max_z = 0.0f;
		if (t < mesh.num_triangles)
		{
			const uint32_t a = mesh.triangles[t * 3], b = mesh.triangles[t * 3 + 1], c = mesh.triangles[t * 3 + 2];
			const V3 v0 = {v[a * 8], v[a * 8 + 1], v[a * 8 + 2]}, v1 = {v[b * 8], v[b * 8 + 1], v[b * 8 + 2]}, v2 = {v[c * 8], v[c * 8 + 1], v[c * 8 + 2]};
			keep = dot3(cross3(v1 - v0, v2 - v1), v0) > 0.0f; // get_triangle_plane (:366-374)
			max_z = fmaxf(fmaxf(v0.z, v1.z), v2.z);
		}
		const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, keep);
		if (lane == 0)
			s_warp[warp] = __popc(ballot);
		__syncthreads();
		uint32_t before = s_base;
		for (int w = 0; w < warp; ++w)
			before += s_warp[w];
		if (keep)
			keys[before + __popc(ballot & ((1u << lane) - 1))] = ((unsigned long long)descending_key(max_z) << 32) | t;
		__syncthreads();
		if (threadIdx.x == 0)
		{
			uint32_t total = 0;
			for (uint32_t w = 0; w < blockDim.x / 32; ++w)
				total += s_warp[w];
			s_base += total;
		}
		__syncthreads();
	}
	const uint32_t kept = s_base;
	uint32_t n = 1;
	while (n < kept)
		n <<= 1;
	for (uint32_t i = kept + threadIdx.x; i < n; i += blockDim.x)
		keys[i] = ~0ull; // padding sorts last
	__syncthreads();
	// bitonic sort, ascending: larger max z first, mesh order among equal keys
	for (uint32_t k = 2; k <= n; k <<= 1)
		for (uint32_t j = k >> 1; j > 0; j >>= 1)
		{
			for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
			{
				const uint32_t partner = i ^ j;
				if (partner > i)
				{
					const unsigned long long a = keys[i], b = keys[partner];
					const bool up = (i & k) == 0;
					if ((a > b) == up)
					{
						keys[i] = b;
						keys[partner] = a;
					}
				}
			}
			__syncthreads();
		}
	uint16_t* out = out_triangles + (size_t)in.triangle_offset * 3;
	for (uint32_t i = threadIdx.x; i < kept; i += blockDim.x)
	{
		const uint32_t t = (uint32_t)keys[i];
		out[i * 3] = mesh.triangles[t * 3];
		out[i * 3 + 1] = mesh.triangles[t * 3 + 1];
		out[i * 3 + 2] = mesh.triangles[t * 3 + 2];
	}
	if (threadIdx.x == 0)
		out_counts[blockIdx.x] = kept;
}

} // namespace

cudaError_t launch_models_prepare(const ModelLaunch& l, cudaStream_t stream)
{
	if (l.num_instances == 0)
		return cudaSuccess;
	if (l.max_bones)
		model_bones_kernel<<<l.num_instances, 256, (size_t)std::min<uint32_t>(l.max_bones, 255u) * 16 * sizeof(float), stream>>>(l.instances, l.meshes, l.bone_scratch);
	model_vertices_kernel<<<dim3((l.max_vertices + 127) / 128, l.num_instances), 128, 0, stream>>>(l.instances, l.meshes, l.bone_scratch, l.vertices);
	model_triangles_kernel<<<l.num_instances, 1024, 0, stream>>>(l.instances, l.meshes, l.vertices, l.sort_scratch, l.triangles, l.counts);
	return cudaGetLastError();
}

} // namespace sqw
