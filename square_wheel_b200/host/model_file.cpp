// Triangle-model file loaders: Quake III MD3 and Inter-Quake Model (IQM v2), the two formats the reference's resources manager
// hands to the renderer (src/square_wheel/triangle_model_md3.rs, triangle_model_iqm.rs, triangle_model_loading.rs).  SURVEY §8 f4.
//
// Both loaders work on a byte image of the file (the reference seeks and reads; a short read there is an io::Error, here it is
// a load failure with a message) and fill the same TriangleModel the seeded fixtures use, so a loaded model goes through the
// unchanged animation / lighting / clipping / sorting mirror (models.cpp) into SQW_CMD_TRIANGLE commands.
#include "models.hpp"
#include <cstdio>
#include <cstring>
#include <string>

namespace sqwh
{
namespace
{

struct Reader
{
	const uint8_t* data;
	size_t size;
	std::string* err;
	bool fail(const char* what) const
	{
		if (err)
			*err = what;
		return false;
	}
	// read_vector / read_chunk (triangle_model_loading.rs:3-30): `count` elements of T at `offset`
	template <class T>
	bool read(uint64_t offset, uint64_t count, std::vector<T>& out, const char* what) const
	{
		out.clear();
		if (count == 0)
			return true;
		if (offset > size || count > (size - offset) / sizeof(T))
			return fail(what);
		out.resize((size_t)count);
		std::memcpy(out.data(), data + offset, (size_t)count * sizeof(T));
		return true;
	}
	template <class T>
	bool read_one(uint64_t offset, T& out, const char* what) const
	{
		if (offset > size || sizeof(T) > size - offset)
			return fail(what);
		std::memcpy(&out, data + offset, sizeof(T));
		return true;
	}
};

std::string fixed_str(const uint8_t* s, size_t n) // get_str (triangle_model_md3.rs:201-205)
{
	size_t l = 0;
	while (l < n && s[l] != 0)
		++l;
	return std::string(reinterpret_cast<const char*>(s), l);
}

// ---- MD3 (triangle_model_md3.rs:207-300) ---------------------------------------------------------------------
constexpr size_t MAX_QPATH = 64;
constexpr uint32_t MD3_VERSION = 15;
constexpr float MD3_COORD_SCALE = 1.0f / 64.0f;

#pragma pack(push, 1)
struct Md3Header
{
	uint8_t ident[4];
	uint32_t version;
	uint8_t name[MAX_QPATH];
	uint32_t flags;
	uint32_t num_frames, num_tags, num_meshes, num_skins;
	uint32_t lump_frameinfo, lump_tags, lump_meshes, lump_end;
};
struct Md3Mesh
{
	uint8_t ident[4];
	uint8_t name[MAX_QPATH];
	uint32_t flags;
	uint32_t num_frames, num_shaders, num_vertices, num_triangles;
	uint32_t lump_triangles, lump_shaders, lump_texcoords, lump_framevertices, lump_end;
};
struct Md3Shader
{
	uint8_t name[MAX_QPATH];
	uint32_t index;
};
struct Md3Tag
{
	uint8_t name[MAX_QPATH];
	float origin[3];
	float rotation_matrix[9];
};
struct Md3Frame
{
	float mins[3], maxs[3], origin[3];
	float radius;
	uint8_t name[16];
};
struct Md3Vertex
{
	int16_t origin[3];
	int16_t normal_pitch_yaw;
};
struct Md3Triangle
{
	uint32_t v[3];
};
struct Md3TexCoord
{
	float tc[2];
};
#pragma pack(pop)
static_assert(sizeof(Md3Header) == 108 && sizeof(Md3Mesh) == 108 && sizeof(Md3Tag) == 112 && sizeof(Md3Frame) == 56 && sizeof(Md3Shader) == 68 && sizeof(Md3Vertex) == 8, "MD3 layout");

Vec3f md3_decompress_normal(int16_t normal_pitch_yaw) // :191-199
{
	const float scale = 6.28318530717958647692f / 256.0f;
	const float pitch = (float)(normal_pitch_yaw & 255) * scale;
	const float yaw = (float)((normal_pitch_yaw >> 8) & 255) * scale;
	const float pitch_sin = std::sin(pitch);
	return {pitch_sin * std::cos(yaw), pitch_sin * std::sin(yaw), std::cos(pitch)};
}

bool load_md3_mesh(const Reader& rd, const Md3Mesh& src, uint64_t mesh_offset, TriangleModelMesh& mesh, bool& skip) // :118-189
{
	skip = false;
	if (std::memcmp(src.ident, "IDP3", 4) != 0)
	{
		skip = true; // "Mesh is not a valid MD3 model": the reference drops the mesh and goes on
		return true;
	}
	std::vector<Md3Triangle> triangles;
	std::vector<Md3TexCoord> tex_coords;
	std::vector<Md3Vertex> frames;
	std::vector<Md3Shader> shaders;
	if (!rd.read(src.lump_triangles + mesh_offset, src.num_triangles, triangles, "MD3: truncated triangles") ||
		!rd.read(src.lump_texcoords + mesh_offset, src.num_vertices, tex_coords, "MD3: truncated texture coordinates") ||
		!rd.read(src.lump_framevertices + mesh_offset, (uint64_t)src.num_vertices * src.num_frames, frames, "MD3: truncated frame vertices") ||
		!rd.read(src.lump_shaders + mesh_offset, src.num_shaders, shaders, "MD3: truncated shaders"))
		return false;

	mesh = TriangleModelMesh();
	mesh.name = fixed_str(src.name, MAX_QPATH);
	mesh.material_name = shaders.empty() ? std::string() : fixed_str(shaders[0].name, MAX_QPATH);
	mesh.num_frames = src.num_frames;
	mesh.triangles.resize(triangles.size());
	for (size_t i = 0; i < triangles.size(); ++i)
		mesh.triangles[i] = {(uint16_t)triangles[i].v[0], (uint16_t)triangles[i].v[1], (uint16_t)triangles[i].v[2]}; // `as VertexIndex`

	std::vector<VertexAnimatedVertexVariable> variable(frames.size());
	for (size_t i = 0; i < frames.size(); ++i)
	{
		variable[i].position = Vec3f{(float)frames[i].origin[0], (float)frames[i].origin[1], (float)frames[i].origin[2]} * MD3_COORD_SCALE;
		variable[i].normal = md3_decompress_normal(frames[i].normal_pitch_yaw);
	}
	if (src.num_frames == 1 && tex_coords.size() == variable.size())
	{
		mesh.kind = VERTEX_DATA_NON_ANIMATED;
		mesh.non_animated.resize(variable.size());
		for (size_t i = 0; i < variable.size(); ++i)
		{
			mesh.non_animated[i].position = variable[i].position;
			mesh.non_animated[i].normal = variable[i].normal;
			mesh.non_animated[i].tex_coord[0] = tex_coords[i].tc[0];
			mesh.non_animated[i].tex_coord[1] = tex_coords[i].tc[1];
		}
	}
	else
	{
		mesh.kind = VERTEX_DATA_VERTEX_ANIMATED;
		mesh.constant.resize(tex_coords.size());
		for (size_t i = 0; i < tex_coords.size(); ++i)
		{
			mesh.constant[i].tex_coord[0] = tex_coords[i].tc[0];
			mesh.constant[i].tex_coord[1] = tex_coords[i].tc[1];
		}
		mesh.variable.swap(variable);
	}
	return true;
}

// ---- IQM (triangle_model_iqm.rs:428-559) ---------------------------------------------------------------------
#pragma pack(push, 1)
struct IQMHeader
{
	uint8_t magic[16];
	uint32_t version, filesize, flags;
	uint32_t num_text, ofs_text;
	uint32_t num_meshes, ofs_meshes;
	uint32_t num_vertexarrays, num_vertexes, ofs_vertexarrays;
	uint32_t num_triangles, ofs_triangles, ofs_adjacency;
	uint32_t num_joints, ofs_joints;
	uint32_t num_poses, ofs_poses;
	uint32_t num_anims, ofs_anims;
	uint32_t num_frames, num_framechannels, ofs_frames, ofs_bounds;
	uint32_t num_comment, ofs_comment;
	uint32_t num_extensions, ofs_extensions;
};
struct IQMMesh
{
	uint32_t name, material, first_vertex, num_vertexes, first_triangle, num_triangles;
};
struct IQMBounds
{
	float bbmins[3], bbmaxs[3];
	float xyradius, radius;
};
struct IQMTriangle
{
	uint32_t vertex[3];
};
struct IQMVertexArray
{
	uint32_t type, flags, format, size, offset;
};
struct IQMJoint
{
	uint32_t name, parent;
	float translate[3], rotate[4], scale[3];
};
struct IQMPose
{
	uint32_t parent, channelmask;
	float channeloffset[10], channelscale[10];
};
struct IQMAnim
{
	uint32_t name, first_frame, num_frames;
	float framerate;
	uint32_t flags;
};
#pragma pack(pop)
static_assert(sizeof(IQMHeader) == 124 && sizeof(IQMJoint) == 48 && sizeof(IQMPose) == 88 && sizeof(IQMBounds) == 32 && sizeof(IQMVertexArray) == 20, "IQM layout");

constexpr uint32_t IQM_VERSION = 2;
constexpr uint32_t IQM_UBYTE = 1, IQM_FLOAT = 7;
constexpr uint32_t IQM_POSITION = 0, IQM_TEXCOORD = 1, IQM_NORMAL = 2, IQM_BLENDINDEXES = 4, IQM_BLENDWEIGHTS = 5;
constexpr uint32_t IQM_LOOP = 1;

std::string iqm_text(const std::vector<uint8_t>& texts, uint32_t offset) // get_text_str (:138-147)
{
	size_t start = offset, end = start;
	if (start > texts.size())
		return std::string();
	while (end < texts.size() && texts[end] != 0)
		++end;
	return std::string(reinterpret_cast<const char*>(texts.data()) + start, end - start);
}

Mat4f iqm_bone_matrix(Vec3f scale, Quaternion rotate, Vec3f translate) // get_bone_matrix (:419-426)
{
	// rotate / rotate.magnitude(): cgmath's magnitude = sqrt(s*s + v.dot(v)), the division is per component
	const float magnitude = std::sqrt(rotate.s * rotate.s + dot(rotate.v, rotate.v));
	const Quaternion unit{rotate.s / magnitude, Vec3f{rotate.v.x / magnitude, rotate.v.y / magnitude, rotate.v.z / magnitude}};
	return Mat4f::from_translation(translate) * Mat4f::from_nonuniform_scale(scale.x, scale.y, scale.z) * mat4_from_quaternion(unit);
}

// create_frames (:294-417): per frame and joint, parent_base * (pose * inverse_base).
bool iqm_create_frames(const Reader& rd, const IQMHeader& header, const std::vector<IQMJoint>& joints, const std::vector<IQMPose>& poses, std::vector<Mat4f>& frame_bones)
{
	std::vector<Mat4f> base(joints.size(), Mat4f::identity()), inverse_base(joints.size(), Mat4f::identity());
	for (size_t i = 0; i < joints.size(); ++i)
	{
		const IQMJoint& j = joints[i];
		const Mat4f mat = iqm_bone_matrix(Vec3f{j.scale[0], j.scale[1], j.scale[2]}, Quaternion{j.rotate[3], Vec3f{j.rotate[0], j.rotate[1], j.rotate[2]}}, Vec3f{j.translate[0], j.translate[1], j.translate[2]});
		Mat4f inverse_mat;
		if (!mat.invert(inverse_mat))
			return rd.fail("IQM: singular joint matrix"); // the reference unwraps (panics) here
		if (j.parent < joints.size())
		{
			// (a parent listed after its child still holds the identity here, as in the reference)
			base[i] = base[j.parent] * mat;
			inverse_base[i] = inverse_mat * inverse_base[j.parent];
		}
		else
		{
			base[i] = mat;
			inverse_base[i] = inverse_mat;
		}
	}

	frame_bones.assign(poses.size() * (size_t)header.num_frames, Mat4f::identity());
	// the rest of the file from ofs_frames on, as u16 channel values
	const size_t frames_begin = std::min<size_t>(header.ofs_frames, rd.size);
	const size_t num_values = (rd.size - frames_begin) / 2;
	size_t pos = 0;
	bool short_data = false;
	auto next = [&]() -> float {
		if (pos >= num_values)
		{
			short_data = true; // the reference indexes out of bounds (panic)
			return 0.0f;
		}
		uint16_t v;
		std::memcpy(&v, rd.data + frames_begin + pos * 2, 2);
		++pos;
		return (float)v;
	};
	for (uint32_t frame = 0; frame < header.num_frames; ++frame)
	{
		Mat4f* dst = frame_bones.data() + (size_t)frame * poses.size();
		for (size_t p = 0; p < poses.size(); ++p)
		{
			const IQMPose& pose = poses[p];
			float c[10];
			for (int k = 0; k < 10; ++k)
			{
				c[k] = pose.channeloffset[k];
				if (pose.channelmask & (1u << k))
					c[k] += next() * pose.channelscale[k];
			}
			const Mat4f mat = iqm_bone_matrix(Vec3f{c[7], c[8], c[9]}, Quaternion{c[6], Vec3f{c[3], c[4], c[5]}}, Vec3f{c[0], c[1], c[2]});
			const Mat4f mat1 = mat * inverse_base[p];
			dst[p] = pose.parent < poses.size() ? base[pose.parent] * mat1 : mat1;
		}
	}
	if (short_data)
		return rd.fail("IQM: truncated frame data");
	return true;
}

} // namespace

bool load_model_md3(const uint8_t* data, size_t size, TriangleModel& out, std::string* err) // triangle_model_md3.rs:5-116
{
	const Reader rd{data, size, err};
	Md3Header header;
	if (!rd.read_one(0, header, "Can't read Md3 header"))
		return false;
	if (std::memcmp(header.ident, "IDP3", 4) != 0)
		return rd.fail("File is not a valid MD3 model");
	if (header.version != MD3_VERSION)
		return rd.fail("Can't load incompatible MD3 model version");

	std::vector<Md3Frame> frames;
	std::vector<Md3Tag> tags;
	if (!rd.read(header.lump_frameinfo, header.num_frames, frames, "MD3: truncated frames") ||
		!rd.read(header.lump_tags, (uint64_t)header.num_tags * header.num_frames, tags, "MD3: truncated tags"))
		return false;

	out = TriangleModel();
	out.frames_info.resize(frames.size());
	for (size_t i = 0; i < frames.size(); ++i)
		out.frames_info[i] = BBox::from_min_max(Vec3f{frames[i].mins[0], frames[i].mins[1], frames[i].mins[2]}, Vec3f{frames[i].maxs[0], frames[i].maxs[1], frames[i].maxs[2]});
	// a tag set exists for every frame: the tags of frame 0 name the bones, every tag is a bone matrix of its frame
	for (size_t i = 0; i < std::min<size_t>(header.num_tags, tags.size()); ++i)
	{
		out.bones.push_back(0xFFFFFFFFu);
		out.bone_names.push_back(fixed_str(tags[i].name, MAX_QPATH));
	}
	out.frame_bones.resize(tags.size());
	for (size_t i = 0; i < tags.size(); ++i)
	{
		const float* r = tags[i].rotation_matrix;
		const Mat4f rotation = {{r[0], r[1], r[2], 0}, {r[3], r[4], r[5], 0}, {r[6], r[7], r[8], 0}, {0, 0, 0, 1}}; // Mat4f::from(Mat3f::new(..)): columns
		out.frame_bones[i] = Mat4f::from_translation(Vec3f{tags[i].origin[0], tags[i].origin[1], tags[i].origin[2]}) * rotation;
	}

	uint64_t offset = header.lump_meshes;
	for (uint32_t i = 0; i < header.num_meshes; ++i)
	{
		Md3Mesh mesh_header;
		std::string ignored;
		const Reader quiet{data, size, &ignored};
		if (!quiet.read_one(offset, mesh_header, "Can't read Md3 mesh header"))
			break; // the reference prints and stops reading meshes
		TriangleModelMesh mesh;
		bool skip = false;
		if (!load_md3_mesh(rd, mesh_header, offset, mesh, skip))
			return false;
		if (!skip)
			out.meshes.push_back(std::move(mesh));
		offset += mesh_header.lump_end;
	}
	out.tc_shift = {-0.5f, -0.5f}; // texture coordinates are floored, not interpolated like Quake III does (:106-107)
	return true;
}

bool load_model_iqm(const uint8_t* data, size_t size, TriangleModel& out, std::string* err) // triangle_model_iqm.rs:5-131
{
	const Reader rd{data, size, err};
	IQMHeader header;
	if (!rd.read_one(0, header, "Can't read IQM header"))
		return false;
	if (std::memcmp(header.magic, "INTERQUAKEMODEL\0", 16) != 0)
		return rd.fail("File is not a valid IQM model");
	if (header.version != IQM_VERSION)
		return rd.fail("Can't load incompatible IQM model version");
	if (header.num_meshes == 0)
		return rd.fail("Invalid IQM model with no meshes");
	if (header.num_joints != header.num_poses)
		return rd.fail("Invalid IQM model: joints/poses mismatch");

	std::vector<uint8_t> texts;
	std::vector<IQMMesh> meshes;
	std::vector<IQMTriangle> triangles;
	std::vector<IQMVertexArray> vertex_arrays;
	std::vector<IQMBounds> bounds;
	std::vector<IQMJoint> joints;
	std::vector<IQMPose> poses;
	std::vector<IQMAnim> anims;
	if (!rd.read(header.ofs_text, header.num_text, texts, "IQM: truncated text") || !rd.read(header.ofs_meshes, header.num_meshes, meshes, "IQM: truncated meshes") ||
		!rd.read(header.ofs_triangles, header.num_triangles, triangles, "IQM: truncated triangles") ||
		!rd.read(header.ofs_vertexarrays, header.num_vertexarrays, vertex_arrays, "IQM: truncated vertex arrays") ||
		!rd.read(header.ofs_bounds, header.num_frames, bounds, "IQM: truncated bounds") || !rd.read(header.ofs_joints, header.num_joints, joints, "IQM: truncated joints") ||
		!rd.read(header.ofs_poses, header.num_poses, poses, "IQM: truncated poses") || !rd.read(header.ofs_anims, header.num_anims, anims, "IQM: truncated animations"))
		return false;
	if (joints.size() > MAX_TRIANGLE_MODEL_BONES)
		return rd.fail("IQM: too many bones");

	// load_vertices (:159-266): only the attribute layouts the reference understands are taken, the rest stays zero
	std::vector<SkeletonAnimatedVertex> vertices(header.num_vertexes, SkeletonAnimatedVertex{{0.0f, 0.0f}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}});
	for (const IQMVertexArray& va : vertex_arrays)
	{
		const size_t n = vertices.size();
		if ((va.type == IQM_POSITION || va.type == IQM_NORMAL) && va.size == 3 && va.format == IQM_FLOAT)
		{
			std::vector<float> v;
			if (!rd.read(va.offset, (uint64_t)n * 3, v, "IQM: truncated vertex array"))
				return false;
			for (size_t i = 0; i < n; ++i)
				(va.type == IQM_POSITION ? vertices[i].position : vertices[i].normal) = Vec3f{v[i * 3], v[i * 3 + 1], v[i * 3 + 2]};
		}
		else if (va.type == IQM_TEXCOORD && va.size == 2 && va.format == IQM_FLOAT)
		{
			std::vector<float> v;
			if (!rd.read(va.offset, (uint64_t)n * 2, v, "IQM: truncated vertex array"))
				return false;
			for (size_t i = 0; i < n; ++i)
				vertices[i].tex_coord[0] = v[i * 2], vertices[i].tex_coord[1] = v[i * 2 + 1];
		}
		else if ((va.type == IQM_BLENDINDEXES || va.type == IQM_BLENDWEIGHTS) && va.size == 4 && va.format == IQM_UBYTE)
		{
			std::vector<uint8_t> v;
			if (!rd.read(va.offset, (uint64_t)n * 4, v, "IQM: truncated vertex array"))
				return false;
			for (size_t i = 0; i < n; ++i)
				std::memcpy(va.type == IQM_BLENDINDEXES ? vertices[i].bone_index : vertices[i].weight, &v[i * 4], 4);
		}
	}

	out = TriangleModel();
	for (const IQMJoint& j : joints)
	{
		out.bones.push_back(j.parent);
		out.bone_names.push_back(iqm_text(texts, j.name));
	}
	if (!iqm_create_frames(rd, header, joints, poses, out.frame_bones))
		return false;
	for (const IQMBounds& b : bounds)
		out.frames_info.push_back(BBox::from_min_max(Vec3f{b.bbmins[0], b.bbmins[1], b.bbmins[2]}, Vec3f{b.bbmaxs[0], b.bbmaxs[1], b.bbmaxs[2]}));
	if (out.frames_info.empty()) // at least one frame for non-animated models (:78-87)
	{
		if (vertices.empty())
			return rd.fail("IQM: model without vertices"); // the reference indexes vertices[0] (panic)
		BBox bbox = BBox::from_point(vertices[0].position);
		for (const SkeletonAnimatedVertex& v : vertices)
			bbox.extend_with_point(v.position);
		out.frames_info.push_back(bbox);
	}
	for (const IQMAnim& a : anims)
		out.animations.push_back(TriangleModelAnimation{iqm_text(texts, a.name), a.first_frame, a.num_frames, a.framerate, (a.flags & IQM_LOOP) != 0});

	// the reference loads ONE mesh holding every triangle and vertex of the file, named after the first mesh (:113-122)
	TriangleModelMesh mesh;
	mesh.name = iqm_text(texts, meshes[0].name);
	mesh.material_name = iqm_text(texts, meshes[0].material);
	mesh.kind = VERTEX_DATA_SKELETON_ANIMATED;
	mesh.num_frames = header.num_frames;
	mesh.triangles.resize(triangles.size());
	for (size_t i = 0; i < triangles.size(); ++i)
		mesh.triangles[i] = {(uint16_t)triangles[i].vertex[0], (uint16_t)triangles[i].vertex[1], (uint16_t)triangles[i].vertex[2]};
	mesh.skeleton.swap(vertices);
	out.meshes.push_back(std::move(mesh));
	out.tc_shift = {-0.5f, -0.5f};
	return true;
}

// resources_manager.rs picks the loader by file extension; a byte image is recognised by its magic.
bool load_model(const uint8_t* data, size_t size, TriangleModel& out, std::string* err)
{
	if (size >= 16 && std::memcmp(data, "INTERQUAKEMODEL\0", 16) == 0)
		return load_model_iqm(data, size, out, err);
	if (size >= 4 && std::memcmp(data, "IDP3", 4) == 0)
		return load_model_md3(data, size, out, err);
	if (err)
		*err = "unknown model format (neither IQM nor MD3)";
	return false;
}

bool load_model_file(const char* path, TriangleModel& out, std::string* err)
{
	std::FILE* f = std::fopen(path, "rb");
	if (!f)
	{
		if (err)
			*err = std::string("cannot open ") + path;
		return false;
	}
	std::vector<uint8_t> bytes;
	uint8_t buf[65536];
	size_t n;
	while ((n = std::fread(buf, 1, sizeof(buf), f)) > 0)
		bytes.insert(bytes.end(), buf, buf + n);
	std::fclose(f);
	return load_model(bytes.data(), bytes.size(), out, err);
}

} // namespace sqwh
