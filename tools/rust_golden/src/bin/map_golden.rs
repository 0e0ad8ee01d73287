// Golden summary of the reference's MAP TOOLS for tests/test_map_compiler_cpu.py (the C++ restatement is
// square_wheel_b200/host/map_compiler.cpp).  Calls only public functions of the unmodified reference, in the order of
// map_compiler/main.rs:38-125:
//   map_file_q1::parse_map_file_content, map_polygonizer::polygonize_map, map_csg::perform_csg_for_map_brushes,
//   bsp_builder::build_leaf_bsp_tree / build_submodel_bsp_tree, bsp_map_compact_conversion::convert_bsp_map_to_compact_format,
//   lightmaps_builder::build_dummy_lightmaps
// and prints ONE JSON object of order-independent facts about the compiled map (the reference iterates HashMaps where the
// restatement keeps file / tree order, so element order is not comparable; counts and sorted multisets are).
//
//   python tests/golden/make_map_fixture.py                # writes tests/golden/two_rooms.map (the fixture of the tests)
//   cd tools/rust_golden && cargo run --release --bin map_golden -- ../../tests/golden/two_rooms.map \
//        > ../../tests/golden/reference_map_compiler.json
// Not built in this repository's image (no rustc).
use square_wheel_lib::common::{
	bsp_builder, bsp_map_compact_conversion, lightmaps_builder, map_csg, map_file_q1, map_polygonizer, material,
};

fn polygon_area(vertices: &[cgmath::Vector3<f32>]) -> f64
{
	use cgmath::InnerSpace;
	let mut sum = cgmath::Vector3::<f64>::new(0.0, 0.0, 0.0);
	for i in 1 .. vertices.len().max(2) - 1
	{
		let a = (vertices[i] - vertices[0]).cast::<f64>().unwrap();
		let b = (vertices[i + 1] - vertices[0]).cast::<f64>().unwrap();
		sum += a.cross(b);
	}
	0.5 * sum.magnitude()
}

fn main()
{
	let args: Vec<String> = std::env::args().collect();
	let text = std::fs::read_to_string(&args[1]).unwrap();
	// the fixture's materials are all defaults (no materials directory): the same table the Python test passes
	let materials = material::MaterialsMap::new();
	let parsed = map_file_q1::parse_map_file_content(&text).unwrap();
	let polygonized = map_polygonizer::polygonize_map(&parsed);
	let initial_polygons: usize = polygonized[0].brushes.iter().map(|b| b.len()).sum();
	let mut out = serde_json::Map::new();
	for (name, perform_csg, advanced) in [("csg_advanced", true, true), ("plain", false, false)]
	{
		let csg = if perform_csg
		{
			map_csg::perform_csg_for_map_brushes(&polygonized, &materials)
		}
		else
		{
			map_csg::perform_no_csg_for_map_brushes(&polygonized)
		};
		let tree = bsp_builder::build_leaf_bsp_tree(&csg, &materials, advanced);
		let submodels = csg[1 ..]
			.iter()
			.map(|s| bsp_builder::build_submodel_bsp_tree(s, &materials))
			.collect::<Vec<_>>();
		let mut map = bsp_map_compact_conversion::convert_bsp_map_to_compact_format(&tree, &csg, &submodels, &materials);
		lightmaps_builder::build_dummy_lightmaps(&materials, &mut map);
		let mut areas: Vec<f64> = map
			.polygons
			.iter()
			.map(|p| polygon_area(&map.vertices[p.first_vertex as usize .. (p.first_vertex + p.num_vertices) as usize]))
			.collect();
		areas.sort_by(|a, b| a.partial_cmp(b).unwrap());
		let mut portal_areas: Vec<f64> = map
			.portals
			.iter()
			.map(|p| polygon_area(&map.vertices[p.first_vertex as usize .. (p.first_vertex + p.num_vertices) as usize]))
			.collect();
		portal_areas.sort_by(|a, b| a.partial_cmp(b).unwrap());
		out.insert(
			name.to_string(),
			serde_json::json!({
				"initial_polygons": initial_polygons,
				"polygons_after_csg": csg[0].polygons.len(),
				"nodes": map.nodes.len(),
				"leafs": map.leafs.len(),
				"polygons": map.polygons.len(),
				"portals": map.portals.len(),
				"submodels": map.submodels.len(),
				"lightmap_texels": map.lightmaps_data.len(),
				"total_polygon_area": areas.iter().sum::<f64>(),
				"total_portal_area": portal_areas.iter().sum::<f64>(),
				"polygon_areas_sorted": areas,
				"portal_areas_sorted": portal_areas,
			}),
		);
	}
	println!("{}", serde_json::Value::Object(out));
}
