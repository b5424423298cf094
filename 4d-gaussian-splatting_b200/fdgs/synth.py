"""Deterministic synthetic scenes for the parity tests and the benchmark.

Everything is generated on the CPU from a seeded torch.Generator (fp32) and moved afterwards, so
the CUDA path, the CPU oracle and the reference rasterizer all see bit-identical inputs.
Camera construction follows the reference (scene/cameras.py:59-71, utils/graphics_utils.py:39-94):
matrices are stored transposed, i.e. column-major in memory, as the kernels expect.
"""
import math
from dataclasses import dataclass, field
from typing import Optional

import torch


def projection_matrix(znear, zfar, fovx, fovy):
    """utils/graphics_utils.py:57-77 getProjectionMatrix (row-major, before the transpose)."""
    tan_y, tan_x = math.tan(fovy / 2), math.tan(fovx / 2)
    top, right = tan_y * znear, tan_x * znear
    bottom, left = -top, -right
    P = torch.zeros(4, 4)
    z_sign = 1.0
    P[0, 0] = 2.0 * znear / (right - left)
    P[1, 1] = 2.0 * znear / (top - bottom)
    P[0, 2] = (right + left) / (right - left)
    P[1, 2] = (top + bottom) / (top - bottom)
    P[3, 2] = z_sign
    P[2, 2] = z_sign * zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


@dataclass
class Camera:
    image_width: int
    image_height: int
    FoVx: float
    FoVy: float
    world_view_transform: torch.Tensor   # [4,4] transposed (column-major in memory)
    full_proj_transform: torch.Tensor
    camera_center: torch.Tensor
    timestamp: float = 0.5

    def to(self, device):
        return Camera(self.image_width, self.image_height, self.FoVx, self.FoVy,
                      self.world_view_transform.to(device), self.full_proj_transform.to(device),
                      self.camera_center.to(device), self.timestamp)


def make_camera(W, H, timestamp=0.5, focal_scale=0.54, negative_fov=False, R=None, T=None):
    """Camera at the origin looking down +z (world_view = I) unless R/T are given.
    negative_fov reproduces the N3V quirk FoVx = FoVy = -1 (dataset_readers.py:275-293)."""
    fl = focal_scale * W
    if negative_fov:
        fovx = fovy = -1.0
    else:
        fovx = 2 * math.atan(W / (2 * fl))
        fovy = 2 * math.atan(H / (2 * fl))
    Rt = torch.eye(4)
    if R is not None:
        Rt[:3, :3] = R.t()
        Rt[:3, 3] = T
    view = Rt.t().contiguous()                       # world_view_transform, cameras.py:65
    proj = projection_matrix(0.01, 100.0, fovx, fovy).t().contiguous()
    full = (view.unsqueeze(0).bmm(proj.unsqueeze(0))).squeeze(0).contiguous()
    center = view.inverse()[3, :3].contiguous()
    return Camera(W, H, fovx, fovy, view, full, center, timestamp)


def pose_from_euler(yaw_deg, pitch_deg, roll_deg, ex, ey, ez):
    """Camera-to-world rotation R (columns = camera axes) from yaw (about y), pitch (about x), roll (about z), and
    the world-to-camera translation T = -R^T eye of a camera at `eye` -- the (R, T) pair make_camera() takes
    (reference convention: scene/cameras.py:59-71 with utils/graphics_utils.py:39-55 getWorld2View2)."""
    y, p, r = (math.radians(a) for a in (yaw_deg, pitch_deg, roll_deg))
    Ry = torch.tensor([[math.cos(y), 0.0, math.sin(y)], [0.0, 1.0, 0.0], [-math.sin(y), 0.0, math.cos(y)]])
    Rx = torch.tensor([[1.0, 0.0, 0.0], [0.0, math.cos(p), -math.sin(p)], [0.0, math.sin(p), math.cos(p)]])
    Rz = torch.tensor([[math.cos(r), -math.sin(r), 0.0], [math.sin(r), math.cos(r), 0.0], [0.0, 0.0, 1.0]])
    R = (Ry @ Rx @ Rz).contiguous()
    eye = torch.tensor([ex, ey, ez])
    return R, -(R.t() @ eye)


@dataclass
class Scene:
    """Activated (post exp/normalize/sigmoid) Gaussian parameters, i.e. the rasterizer's inputs."""
    means3D: torch.Tensor
    ts: torch.Tensor
    scales: torch.Tensor
    scales_t: torch.Tensor
    rotations: torch.Tensor
    rotations_r: torch.Tensor
    opacities: torch.Tensor
    shs: torch.Tensor            # [P, M, 3]
    flow_2d: torch.Tensor        # [P, 2]
    sh_degree: int = 3
    sh_degree_t: int = 2
    time_duration: float = 1.0
    rot_4d: bool = True
    gaussian_dim: int = 4
    force_sh_3d: bool = False
    extras: dict = field(default_factory=dict)

    @property
    def P(self):
        return self.means3D.shape[0]

    def tensors(self):
        return dict(means3D=self.means3D, ts=self.ts, scales=self.scales, scales_t=self.scales_t,
                    rotations=self.rotations, rotations_r=self.rotations_r, opacities=self.opacities,
                    shs=self.shs, flow_2d=self.flow_2d)

    def to(self, device):
        kw = {k: v.to(device) for k, v in self.tensors().items()}
        return Scene(**kw, sh_degree=self.sh_degree, sh_degree_t=self.sh_degree_t, time_duration=self.time_duration,
                     rot_4d=self.rot_4d, gaussian_dim=self.gaussian_dim, force_sh_3d=self.force_sh_3d,
                     extras=dict(self.extras))


def make_scene(P, cam: Camera, seed, M=48, sh_degree=3, sh_degree_t=2, sigma_px=1.46016, flow=False,
               zmin=2.0, zmax=10.0, frustum=1.1, scale_t_mean=0.15, time_duration=1.0,
               rot_4d=True, gaussian_dim=4, force_sh_3d=False, opacity_lo=0.1, opacity_hi=0.9) -> Scene:
    """SURVEY.md section 8(d) synthetic inputs: z~U(zmin,zmax), x,y uniform in `frustum` x the view
    frustum, t~U(0,1), log-scales ~ N(log(sigma_px * z / focal), 0.6^2) -- which is exactly
    log(0.002 z) at the benchmark camera (W=1352, focal=0.54 W) and keeps the same on-screen
    footprint (~1.5 px sigma) at the small test resolutions --, unit random quaternions,
    opacity ~ U(0.1,0.9), SH dc = (U-0.5)/C0, rest ~ N(0, 0.05^2)."""
    g = torch.Generator().manual_seed(seed)
    tanx = abs(math.tan(cam.FoVx * 0.5))
    tany = abs(math.tan(cam.FoVy * 0.5))
    u = lambda *s: torch.rand(*s, generator=g)
    n = lambda *s: torch.randn(*s, generator=g)
    z = zmin + (zmax - zmin) * u(P)
    x = (2 * u(P) - 1) * frustum * tanx * z
    y = (2 * u(P) - 1) * frustum * tany * z
    means = torch.stack([x, y, z], dim=1)
    # the points are drawn in the camera frame; a posed camera sees them through world = R (p_cam - T)
    w2c = cam.world_view_transform.t()            # [4,4] row-major world-to-camera
    Rc, Tc = w2c[:3, :3], w2c[:3, 3]              # p_cam = Rc p_world + Tc
    if not (torch.equal(Rc, torch.eye(3)) and torch.equal(Tc, torch.zeros(3))):
        means = (means - Tc) @ Rc                 # = Rc^T (p_cam - Tc), row-vector form
    means = means.contiguous()
    ts = u(P, 1) * time_duration
    # ~sigma_px pixels on screen: focal ~= W / (2 tanx)
    px_world = (2 * tanx / cam.image_width) * z
    log_s = torch.log(px_world * sigma_px)[:, None] + 0.6 * n(P, 3)
    scales = torch.exp(log_s).contiguous()
    scales_t = torch.exp(math.log(scale_t_mean) + 0.5 * n(P, 1)).contiguous()
    rot = torch.nn.functional.normalize(n(P, 4), dim=1).contiguous()
    rot_r = torch.nn.functional.normalize(n(P, 4), dim=1).contiguous()
    opac = (opacity_lo + (opacity_hi - opacity_lo) * u(P, 1)).contiguous()
    shs = 0.05 * n(P, M, 3)
    shs[:, 0, :] = (u(P, 3) - 0.5) / 0.28209479177387814
    flow_2d = (0.5 * n(P, 2)) if flow else torch.zeros(P, 2)
    return Scene(means, ts, scales, scales_t, rot, rot_r, opac, shs.contiguous(), flow_2d.contiguous(),
                 sh_degree=sh_degree, sh_degree_t=sh_degree_t, time_duration=time_duration, rot_4d=rot_4d,
                 gaussian_dim=gaussian_dim, force_sh_3d=force_sh_3d)


def raster_settings(cam: Camera, scene: Scene, bg=None, scale_modifier=1.0, debug=False, device=None):
    """GaussianRasterizationSettings field values as render() builds them
    (reference: gaussian_renderer/__init__.py:34-56)."""
    dev = device if device is not None else cam.world_view_transform.device
    if bg is None:
        bg = torch.zeros(3)
    return dict(
        image_height=int(cam.image_height), image_width=int(cam.image_width),
        tanfovx=math.tan(cam.FoVx * 0.5), tanfovy=math.tan(cam.FoVy * 0.5),
        bg=bg.to(dev), scale_modifier=scale_modifier,
        viewmatrix=cam.world_view_transform.to(dev), projmatrix=cam.full_proj_transform.to(dev),
        sh_degree=scene.sh_degree, sh_degree_t=scene.sh_degree_t, campos=cam.camera_center.to(dev),
        timestamp=cam.timestamp, time_duration=scene.time_duration, rot_4d=scene.rot_4d,
        gaussian_dim=scene.gaussian_dim, force_sh_3d=scene.force_sh_3d, prefiltered=False, debug=debug)
