"""Reference-facing mirror of the renderer API (the five entry points that "stay the same"):

    Renderer.new / prepare_frame / draw_frame               src/square_wheel/renderer.rs:24, :61, :106
    Postprocessor.get_hdr_buffer / perform_postprocessing   src/square_wheel/postprocessor.rs:56, :68

The host work (visibility, surface preparation, BSP traversal) runs in the C++ host library; the three pixel stages run in
libsqw_cuda.so through the C-ABI.  The reference's toolchain (Rust) is absent from this image, so this Python layer stands
where `Host::process_frame` (host.rs:254) would call the Rust methods; INTEGRATION.md shows the Rust binding itself.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import _native as N
from .device import Context, Frame, View, make_post_params
from ._native import STAGE_SHADOW_MAPS as N_STAGE_SHADOW_MAPS, STAGE_SURFACES as N_STAGE_SURFACES
from .scene import BACKGROUND_COLOR32, BACKGROUND_COLOR64, FOV_DEFAULT, HostFrame, HostRenderer, Scene


@dataclass
class SurfaceInfo:
    """common/system_window.rs:11-16"""

    width: int
    height: int
    pitch: int = 0

    def __post_init__(self):
        if not self.pitch:
            self.pitch = self.width


@dataclass
class FrameInfo:
    """The part of frame_info.rs:7 the world-drawing path reads: camera (position, azimuth, elevation), FOV, colour modulate."""

    pose: np.ndarray
    fov: float = FOV_DEFAULT
    color_modulate: tuple = (1.0, 1.0, 1.0)
    lights: Optional[np.ndarray] = None  # FrameInfo::lights (frame_info.rs:24) as scene.dynamic_light_dtype records
    model_entities: Optional[np.ndarray] = None  # FrameInfo::model_entities (frame_info.rs:14) as scene.model_entity_dtype records
    sprites: Optional[np.ndarray] = None  # FrameInfo::sprites (frame_info.rs:17) as scene.sprite_dtype records
    decals: Optional[np.ndarray] = None  # FrameInfo::decals (frame_info.rs:16) as scene.decal_dtype records
    submodel_entities: Optional[np.ndarray] = None  # FrameInfo::submodel_entities (frame_info.rs:13) as scene.submodel_entity_dtype records
    portals: Optional[np.ndarray] = None  # FrameInfo::portals (frame_info.rs:18) as scene.portal_dtype records


@dataclass
class RendererConfig:
    """renderer_config.rs:5-31"""

    use_directional_lightmaps: bool = True
    textures_mip_bias: float = 0.0
    clear_background: bool = True
    shadows_quality: float = 0.0
    portals_depth: int = 1  # nested renderers for portals / mirrors (renderer.rs:33: at most 8)
    # The reference never clears a portal texture: texels no nested polygon covers keep whatever an earlier frame left there.
    # True makes them deterministic zeros (what the parity tests compare against).
    clear_portal_textures: bool = False


class Renderer:
    def __init__(self, scene: Scene, config: Optional[RendererConfig] = None, device: int = 0, stream: int = 0):
        """Renderer::new (renderer.rs:24): uploads lightmaps and every texture mip to the GPU (seam S0)."""
        self.scene = scene
        self.config = config or RendererConfig()
        self.ctx = Context(scene.lightmaps, scene.directional_lightmaps, device=device, stream=stream)
        scene.upload(self.ctx)
        self.host = HostRenderer(scene, self.config.use_directional_lightmaps, self.config.textures_mip_bias, self.config.clear_background, self.config.shadows_quality, self.config.portals_depth)
        self._views = {}
        self._frames = {}
        self._portal_views = {}  # (depth, slot) -> (View, Frame): render targets of the nested portal / mirror views
        self._host_frame: Optional[HostFrame] = None
        self._view: Optional[View] = None

    def _get_view(self, surface_info: SurfaceInfo, color_bits: int) -> View:
        key = (surface_info.width, surface_info.height, color_bits)
        if key not in self._views:
            self._views[key] = self.ctx.create_view(surface_info.width, surface_info.height, color_bits)
        return self._views[key]

    def prepare_frame(self, surface_info: SurfaceInfo, frame_info: FrameInfo, color_bits: int = 64):
        """prepare_frame::<ColorT> (renderer.rs:61): host preparation, then surface building on the GPU (seam S1)."""
        hf = self.host.build_frame(surface_info.width, surface_info.height, frame_info.pose, frame_info.fov, frame_info.lights, frame_info.model_entities, color_bits, frame_info.sprites, frame_info.decals, frame_info.submodel_entities, frame_info.portals)
        view = self._get_view(surface_info, color_bits)
        key = (surface_info.width, surface_info.height, color_bits)
        if key not in self._frames:
            self._frames[key] = view.create_frame()
        frame = self._frames[key]
        has_lights = hf.lights is not None and len(hf.lights) > 0
        frame.set_surfaces(hf.jobs, hf.lights if has_lights else None, None, hf.cam_pos, hf.total_surface_pixels, device_shadow_maps_len=hf.shadow_maps_len if has_lights else 0)
        # build_shadow_maps (partial_renderer.rs:194); an empty list clears what an earlier frame recorded
        frame.set_shadow_map_draw(hf.depth_maps, hf.depth_cmds, hf.depth_points)
        frame.submit(N_STAGE_SHADOW_MAPS | N_STAGE_SURFACES)
        self._render_portals(hf, 1, color_bits)  # build_view_portals_textures (partial_renderer.rs:232)
        self._host_frame, self._view, self._frame = hf, view, frame
        return hf

    def _render_portals(self, hf: HostFrame, depth: int, color_bits: int):
        """build_view_portals_textures (partial_renderer.rs:2008-2282): every visible portal gets its own render target; the nested
        frame is built and drawn into it (black at the depth limit), the portal's overlay texture is mixed in, and the polygon
        commands of `hf` that show the portal are pointed at that target (SQW_TEX_VIEW)."""
        for slot, node in enumerate(hf.portals):
            w, h = node.resolution
            key = (depth, slot, color_bits)
            if key in self._portal_views and self._portal_views[key][2] < w * h:  # grown beyond what the target was allocated for
                old_view, old_frame, _ = self._portal_views.pop(key)
                old_frame.close()
                old_view.close()
            if key not in self._portal_views:
                v = self.ctx.create_view(w, (h * 3 + 1) // 2, color_bits)  # sqw_view_set_size works within the allocated AREA
                self._portal_views[key] = (v, v.create_frame(), w * ((h * 3 + 1) // 2))
            view, frame, _ = self._portal_views[key]
            if (view.width, view.height) != (w, h):
                view.set_size(w, h)
            child = node.frame
            if child is None:
                view.draw(np.zeros(0, N.draw_cmd_dtype), None, None, clear_color=0)
            else:
                self._render_portals(child, depth + 1, color_bits)
                has_lights = child.lights is not None and len(child.lights) > 0
                frame.set_surfaces(child.jobs, child.lights if has_lights else None, None, child.cam_pos, child.total_surface_pixels, device_shadow_maps_len=child.shadow_maps_len if has_lights else 0)
                frame.set_shadow_map_draw(child.depth_maps, child.depth_cmds, child.depth_points)
                frame.submit(N_STAGE_SHADOW_MAPS | N_STAGE_SURFACES)
                clear = 0 if self.config.clear_portal_textures else None
                if child.needs_clear:
                    clear = BACKGROUND_COLOR64 if color_bits == 64 else BACKGROUND_COLOR32
                view.draw(child.cmds, child.polygon_points, child.triangle_points, clear_color=clear)
            if node.overlay is not None:
                image, blend, light = node.overlay
                view.mix_texture(image, node.tc_min, blend, light)
            node.view = view
            portal_cmds = (hf.cmds["tex_kind"] == N.TEX_VIEW) & (hf.cmds["kind"] == N.CMD_POLYGON) & (hf.cmds["tex_ref"] == node.index)
            hf.cmds["tex_ref"][portal_cmds] = view.index

    def draw_frame(self, pixels: Optional[np.ndarray] = None):
        """draw_frame::<ColorT> (renderer.rs:106): rasterisation on the GPU (seam S2).  `pixels` ([h, pitch] uint32/uint64), when
        given, receives the picture like the reference's `pixels: &mut [ColorT]`; otherwise it stays on the device."""
        hf, view = self._host_frame, self._view
        clear = None
        if hf.needs_clear:  # draw_background (partial_renderer.rs:249-256)
            clear = BACKGROUND_COLOR64 if view.color_bits == 64 else BACKGROUND_COLOR32
        view.draw(hf.cmds, hf.polygon_points, hf.triangle_points, clear_color=clear)
        if pixels is not None:
            pixels[:, : view.width] = view.read_pixels()
        return view

    def close(self):
        for f in self._frames.values():
            f.close()
        for v, f, _ in self._portal_views.values():
            f.close()
            v.close()
        for v in self._views.values():
            v.close()
        self.host.close()
        self.ctx.close()


@dataclass
class PostprocessorConfig:
    """postprocessor_config.rs:5-93 with the clamps of synchronize_config (postprocessor.rs:958-977)."""

    hdr_rendering: bool = True
    bloom_sigma: float = 1.0
    bloom_buffer_scale_log2: int = 1
    bloom_scale: float = 0.25
    linear_bloom_filter: bool = False
    exposure_update_speed: float = 2.0
    base_brightness: float = 1.0
    zero_level_brightness: float = 1.0
    brightness_scale_power: float = 0.25
    min_exposure: float = 1.0 / 16.0
    max_exposure: float = 16.0
    auto_exposure: bool = True


class Postprocessor:
    def __init__(self, renderer: Renderer, config: Optional[PostprocessorConfig] = None):
        self.renderer = renderer
        self.config = config or PostprocessorConfig()
        self.current_exposure = 1.0  # postprocessor.rs:52
        self._view: Optional[View] = None

    def get_hdr_buffer(self, size) -> View:
        """get_hdr_buffer (postprocessor.rs:56): the 64-bit render target lives on the device."""
        self._view = self.renderer._get_view(SurfaceInfo(size[0], size[1]), 64)
        return self._view

    def perform_postprocessing(self, pixels: Optional[np.ndarray], surface_info: SurfaceInfo, color_modulate, frame_duration_s: float):
        """perform_postprocessing (postprocessor.rs:68): tonemap (+bloom) on the GPU (seam S3); exposure update stays on the host."""
        view = self._view or self.get_hdr_buffer((surface_info.width, surface_info.height))
        c = self.config
        params = make_post_params(self.current_exposure, color_modulate, c.bloom_sigma, c.bloom_scale, c.bloom_buffer_scale_log2, c.linear_bloom_filter)
        out, avg = view.postprocess(params, read_back=pixels is not None)
        if pixels is not None:
            pixels[:, : surface_info.width] = out
        if c.auto_exposure:
            self.update_exposure(avg, frame_duration_s)
        return avg

    def update_exposure(self, average_color_bgr, frame_duration_s: float):
        """update_exposure (postprocessor.rs:206-223) — scalar state for the NEXT frame, stays on the host."""
        c = self.config
        f = np.float32
        brightness = f(min(max(f(f(average_color_bgr[0] + average_color_bgr[1] + average_color_bgr[2]) * f(0.33)), f(1.0 / 1024.0)), f(65536.0))) / f(255.0)
        target = f(math.pow(brightness / f(c.zero_level_brightness), c.brightness_scale_power)) * f(0.5) * f(c.base_brightness) / brightness
        clamped = min(max(float(target), c.min_exposure), c.max_exposure)
        mix = math.exp(-c.exposure_update_speed * frame_duration_s)
        self.current_exposure = float(math.exp(math.log(self.current_exposure) * mix + math.log(clamped) * (1.0 - mix)))
