"""Stand-in for the reference's GaussianModel (scene/gaussian_model.py) holding RAW parameters: the attribute names,
activations and getters render() reads (`_xyz`, `_scaling`, ..., `scaling_activation = torch.exp`, `get_features` =
cat(features_dc, features_rest)) -- enough for both the getter path and the fused raw-parameter entry.
TEST / BENCH INFRASTRUCTURE."""
import torch


class Pipe:
    convert_SHs_python = False
    compute_cov3D_python = False
    debug = False
    env_map_res = 0
    fused_prologue = True


class PipeUnfused(Pipe):
    fused_prologue = False


class RawModel:
    def __init__(self, sc, seed=0, requires_grad=False):
        """Raw parameters whose activations reproduce the activated scene `sc` (up to rounding): log-scales, logits,
        quaternions scaled by random positive lengths, SH rows split into dc / rest."""
        g = torch.Generator().manual_seed(seed)
        dev = sc.means3D.device
        leaf = lambda t: t.detach().clone().contiguous().requires_grad_(requires_grad)
        lens = lambda: (0.5 + torch.rand(sc.P, 1, generator=g)).to(dev)
        self._xyz = leaf(sc.means3D)
        self._t = leaf(sc.ts)
        self._scaling = leaf(torch.log(sc.scales))
        self._scaling_t = leaf(torch.log(sc.scales_t))
        self._rotation = leaf(sc.rotations * lens())
        self._rotation_r = leaf(sc.rotations_r * lens())
        self._opacity = leaf(torch.logit(sc.opacities))
        self._features_dc = leaf(sc.shs[:, :1])
        self._features_rest = leaf(sc.shs[:, 1:])
        self.scaling_activation = torch.exp
        self.opacity_activation = torch.sigmoid
        self.rotation_activation = torch.nn.functional.normalize
        self.active_sh_degree, self.active_sh_degree_t = sc.sh_degree, sc.sh_degree_t
        self.time_duration = [0.0, sc.time_duration]
        self.rot_4d, self.gaussian_dim, self.force_sh_3d = sc.rot_4d, sc.gaussian_dim, sc.force_sh_3d
        self.prefilter_var = -1.0
        self.get_max_sh_channels = sc.shs.shape[1]
        self.env_map = None

    get_xyz = property(lambda s: s._xyz)
    get_t = property(lambda s: s._t)
    get_scaling = property(lambda s: s.scaling_activation(s._scaling))
    get_scaling_t = property(lambda s: s.scaling_activation(s._scaling_t))
    get_rotation = property(lambda s: s.rotation_activation(s._rotation))
    get_rotation_r = property(lambda s: s.rotation_activation(s._rotation_r))
    get_opacity = property(lambda s: s.opacity_activation(s._opacity))
    get_features = property(lambda s: torch.cat((s._features_dc, s._features_rest), dim=1))

    def leaves(self):
        return dict(xyz=self._xyz, t=self._t, scaling=self._scaling, scaling_t=self._scaling_t, rotation=self._rotation,
                    rotation_r=self._rotation_r, opacity=self._opacity, features_dc=self._features_dc,
                    features_rest=self._features_rest)
