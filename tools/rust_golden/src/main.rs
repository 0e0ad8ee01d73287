// Golden-vector emitter for the world-drawing hot path of Square-Wheel (see Cargo.toml and
// tests/golden/make_reference_inputs.py for the file layouts).  Calls only PUBLIC functions of the reference:
//   surfaces::build_surface_simple_lightmap / build_surface_directional_lightmap   (surfaces.rs:66, :95)
//   Rasterizer::new / fill_polygon / fill_triangle                                  (rasterizer.rs:25, :35, :731)
//   Postprocessor::new / get_hdr_buffer / perform_postprocessing                    (postprocessor.rs:39, :56, :68)
use cgmath::Vector3;
use square_wheel_lib::common::{bsp_map_compact::*, color::*, material::BlendingMode, system_window::SurfaceInfo};
use square_wheel_lib::square_wheel::{
	config, debug_stats_printer::DebugStatsPrinter, equations::*, postprocessor::Postprocessor, rasterizer::*, surfaces::*, textures::*,
};
use std::convert::TryInto;

struct Reader<'a>
{
	d: &'a [u8],
	p: usize,
}
impl<'a> Reader<'a>
{
	fn u32(&mut self) -> u32
	{
		let v = u32::from_le_bytes(self.d[self.p .. self.p + 4].try_into().unwrap());
		self.p += 4;
		v
	}
	fn i32(&mut self) -> i32
	{
		self.u32() as i32
	}
	fn f32(&mut self) -> f32
	{
		f32::from_bits(self.u32())
	}
	fn u64(&mut self) -> u64
	{
		let v = u64::from_le_bytes(self.d[self.p .. self.p + 8].try_into().unwrap());
		self.p += 8;
		v
	}
	fn vec3(&mut self) -> Vector3<f32>
	{
		let x = self.f32();
		let y = self.f32();
		let z = self.f32();
		Vector3::new(x, y, z)
	}
}

fn blending(b: u32) -> BlendingMode
{
	match b
	{
		0 => BlendingMode::None,
		1 => BlendingMode::Average,
		2 => BlendingMode::Additive,
		3 => BlendingMode::AlphaTest,
		_ => BlendingMode::AlphaBlend,
	}
}

fn put(out: &mut Vec<u8>, pixels: &[u64])
{
	out.extend_from_slice(&(pixels.len() as u32).to_le_bytes());
	for p in pixels
	{
		out.extend_from_slice(&p.to_le_bytes());
	}
}

fn surface<ColorT: square_wheel_lib::square_wheel::abstract_color::AbstractColor>(
	r: &mut Reader, directional: bool, to_raw: fn(&ColorT) -> u64, zero: ColorT,
) -> Vec<u64>
{
	let basis = PolygonBasisVecs { u: r.vec3(), v: r.vec3(), start: r.vec3(), normal: r.vec3() };
	let surface_size = [r.u32(), r.u32()];
	let tc_min = [r.i32(), r.i32()];
	let tex_size = [r.u32(), r.u32()];
	let flags = r.u32();
	let mut pixels = Vec::with_capacity((tex_size[0] * tex_size[1]) as usize);
	for _ in 0 .. tex_size[0] * tex_size[1]
	{
		// TextureElement is #[repr(C, align(8))] { Color32, PackedNormalRoughness(i32) } = 8 bytes (textures.rs:20-27)
		pixels.push(unsafe { std::mem::transmute::<u64, TextureElement>(r.u64()) });
	}
	let texture = Texture {
		size: tex_size,
		pixels,
		has_normal_map: flags & 1 != 0,
		has_non_one_roughness: flags & 2 != 0,
		is_metal: flags & 4 != 0,
	};
	let lm_size = [r.u32(), r.u32()];
	let lm_scale_log2 = r.u32();
	let lm_shift = [r.u32(), r.u32()];
	let has_lm = r.u32() != 0;
	let n_lm = if has_lm { (lm_size[0] * lm_size[1]) as usize } else { 0 };
	let mut out = vec![zero; (surface_size[0] * surface_size[1]) as usize];
	if directional
	{
		let mut lm = Vec::with_capacity(n_lm);
		for _ in 0 .. n_lm
		{
			let ambient_light = [r.f32(), r.f32(), r.f32()];
			let light_direction_vector_scaled = r.vec3();
			let directional_light_deviation = r.f32();
			let directional_light_color = [r.f32(), r.f32(), r.f32()];
			lm.push(DirectionalLightmapElement { ambient_light, light_direction_vector_scaled, directional_light_deviation, directional_light_color });
		}
		let cam = r.vec3();
		// a polygon without lightmap is black in the reference (partial_renderer.rs:1757-1765 passes an empty slice)
		build_surface_directional_lightmap(&basis, surface_size, tc_min, &texture, lm_size, lm_scale_log2, lm_shift, &lm, &[], &cam, &mut out);
	}
	else
	{
		let mut lm: Vec<LightmapElement> = Vec::with_capacity(n_lm);
		for _ in 0 .. n_lm
		{
			lm.push([r.f32(), r.f32(), r.f32()]);
		}
		let cam = r.vec3();
		build_surface_simple_lightmap(&basis, surface_size, tc_min, &texture, lm_size, lm_scale_log2, lm_shift, &lm, &[], &cam, &mut out);
	}
	out.iter().map(to_raw).collect()
}

fn draw<ColorT: square_wheel_lib::square_wheel::abstract_color::AbstractColor>(
	r: &mut Reader, width: usize, height: usize, clear: ColorT, from_raw: fn(u64) -> ColorT, to_raw: fn(&ColorT) -> u64,
) -> Vec<u64>
{
	let mut fb = vec![clear; width * height];
	let n_cmds = r.u32();
	let info = SurfaceInfo { width, height, pitch: width };
	for _ in 0 .. n_cmds
	{
		let kind = r.u32();
		let blend = blending(r.u32());
		let tc_mode = r.u32();
		let nv = r.u32() as usize;
		let clip = ClipRect { min_x: 0, min_y: 0, max_x: width as i32, max_y: height as i32 };
		if kind == 0
		{
			let mut v = Vec::with_capacity(nv);
			for _ in 0 .. nv
			{
				v.push(PolygonPointProjected { x: r.i32(), y: r.i32() });
			}
			let depth = DepthEquation { d_inv_z_dx: r.f32(), d_inv_z_dy: r.f32(), k: r.f32() };
			let tc = TexCoordEquation { d_tc_dx: [r.f32(), r.f32()], d_tc_dy: [r.f32(), r.f32()], k: [r.f32(), r.f32()] };
			let size = [r.i32(), r.i32()];
			let tex: Vec<ColorT> = (0 .. size[0] * size[1]).map(|_| from_raw(r.u64())).collect();
			let mode = match tc_mode
			{
				0 => TetureCoordinatesInterpolationMode::FullPerspective,
				1 => TetureCoordinatesInterpolationMode::LineZCorrection,
				_ => TetureCoordinatesInterpolationMode::Affine,
			};
			let mut rast = Rasterizer::new(&mut fb, &info, clip);
			rast.fill_polygon(&v, &depth, &tc, &TextureInfo { size }, &tex, mode, blend);
		}
		else
		{
			let mut v = Vec::with_capacity(3);
			for _ in 0 .. 3
			{
				v.push(TrianglePointProjected { x: r.i32(), y: r.i32(), tc: [r.i32(), r.i32()], light: [r.i32(), r.i32(), r.i32()] });
			}
			let size = [r.i32(), r.i32()];
			let tex: Vec<Color32> = (0 .. size[0] * size[1]).map(|_| Color32::from_raw(r.u64() as u32)).collect();
			let mut rast = Rasterizer::new(&mut fb, &info, clip);
			rast.fill_triangle(&[v[0], v[1], v[2]], &TextureInfo { size }, &tex, blend);
		}
	}
	fb.iter().map(to_raw).collect()
}

fn main()
{
	let args: Vec<String> = std::env::args().collect();
	if args.len() != 3
	{
		eprintln!("usage: sqw_rust_golden <reference_inputs.bin> <reference_outputs.bin>");
		std::process::exit(2);
	}
	let data = std::fs::read(&args[1]).expect("inputs");
	let mut r = Reader { d: &data, p: 0 };
	assert_eq!(&data[0 .. 4], b"SQWG");
	r.p = 4;
	assert_eq!(r.u32(), 1);
	let mut out: Vec<u8> = Vec::new();
	out.extend_from_slice(b"SQWO");
	out.extend_from_slice(&1u32.to_le_bytes());
	// which branch of fast_math.rs was compiled in (fast_math.rs:6 / :729)
	let flags: u32 = (cfg!(all(target_arch = "x86_64", target_feature = "sse4.1")) as u32) | ((cfg!(all(target_arch = "x86_64", target_feature = "fma")) as u32) << 1);
	out.extend_from_slice(&flags.to_le_bytes());

	let n_surfaces = r.u32();
	out.extend_from_slice(&n_surfaces.to_le_bytes());
	for _ in 0 .. n_surfaces
	{
		let color_bits = r.u32();
		let directional = r.u32() != 0;
		let px = if color_bits == 32
		{
			surface::<Color32>(&mut r, directional, |c| c.get_raw() as u64, Color32::black())
		}
		else
		{
			surface::<Color64>(&mut r, directional, |c| c.get_raw(), Color64::black())
		};
		put(&mut out, &px);
	}

	let n_draws = r.u32();
	out.extend_from_slice(&n_draws.to_le_bytes());
	for _ in 0 .. n_draws
	{
		let color_bits = r.u32();
		let (w, h) = (r.u32() as usize, r.u32() as usize);
		let clear = r.u64();
		let px = if color_bits == 32
		{
			draw::<Color32>(&mut r, w, h, Color32::from_raw(clear as u32), |x| Color32::from_raw(x as u32), |c| c.get_raw() as u64)
		}
		else
		{
			draw::<Color64>(&mut r, w, h, Color64::from_raw(clear), Color64::from_raw, |c| c.get_raw())
		};
		put(&mut out, &px);
	}

	let n_posts = r.u32();
	out.extend_from_slice(&n_posts.to_le_bytes());
	for _ in 0 .. n_posts
	{
		let (w, h) = (r.u32() as usize, r.u32() as usize);
		let _exposure = r.f32(); // Postprocessor::new starts at 1.0 and a frame duration of 0 keeps it there (postprocessor.rs:206-223)
		let modulate = [r.f32(), r.f32(), r.f32()];
		let bloom_sigma = r.f32();
		let bloom_scale = r.f32();
		let scale_log2 = r.u32();
		let linear = r.u32() != 0;
		let cfg = serde_json::json!({ "postprocessor": {
			"hdr_rendering": true, "use_multithreading": false, "bloom_sigma": bloom_sigma, "bloom_scale": bloom_scale,
			"bloom_buffer_scale_log2": scale_log2, "linear_bloom_filter": linear } });
		let mut post = Postprocessor::new(config::make_shared(cfg));
		{
			let hdr = post.get_hdr_buffer([w, h]);
			for p in hdr.iter_mut()
			{
				*p = Color64::from_raw(r.u64());
			}
		}
		let mut ldr = vec![Color32::black(); w * h];
		let mut printer = DebugStatsPrinter::new(false);
		post.perform_postprocessing(&mut ldr, &SurfaceInfo { width: w, height: h, pitch: w }, &modulate, 0.0, &mut printer);
		let px: Vec<u64> = ldr.iter().map(|c| c.get_raw() as u64).collect();
		put(&mut out, &px);
	}
	assert_eq!(r.p, data.len(), "trailing input");
	std::fs::write(&args[2], &out).expect("outputs");
	println!("{} surfaces, {} draws, {} posts -> {} bytes (sse4.1 {}, fma {})", n_surfaces, n_draws, n_posts, out.len(), flags & 1, (flags >> 1) & 1);
}
