"""Host-side mirror of the reference's networks (rsl_rl/modules/actor_critic.py,
support_networks.py) over libb200gym.so's fused Linear kernels.

`ActorCritic` and `MlpEstimator` keep the reference constructor arguments, method names
(`act`, `act_inference`, `evaluate`, `get_actions_log_prob`, `privileged_encoder`,
`adaptation_encoder`, `action_mean`, `action_std`, `entropy`, `std`) and `state_dict()` keys
and shapes, so checkpoints move both ways.  Parameters live in flat fp32
buffers (one per optimiser: main / adaptation / estimator) so that the gradient all-reduce, the
global-norm clip and Adam are single launches.

Kernel layouts that differ from the checkpoint layout (converted in state_dict / load_state_dict):
  * weights are stored [out, ld] with ld = in rounded up to a multiple of 4 floats (zero padded);
  * AdaptationEncoder: the per-step projection is kept as [B, 10, 32] (30 channels + 2 zero pads), so
    Conv1d(30,20,k=4,s=2) is a GEMM over 4 contiguous steps (K = 4*32, weight [20, k*32+ci]),
    Conv1d(20,10,k=2) a GEMM with K = 2*20 (weight [10, k*20+ci]) writing [B, 3, 12], and
    fc_final reads that time-major tensor (weight [20, t*12+c] instead of Flatten's [20, c*3+t]).
"""
import ctypes as C
import math
from collections import OrderedDict

import torch

from . import _lib


def ceil4(x):
    return (x + 3) // 4 * 4


class FlatGroup:
    """One optimiser's parameters: flat params / grads / exp_avg / exp_avg_sq + device Adam state."""

    def __init__(self):
        self.items = OrderedDict()      # name -> (offset, rows, cols, ld)
        self.size = 0
        self.params = None

    def add(self, name, rows, cols):
        ld = ceil4(cols) if name.endswith(".weight") else cols      # vectors (bias, std) stay contiguous
        self.size = ceil4(self.size)
        self.items[name] = (self.size, rows, cols, ld)
        self.size += rows * ld

    def finalize(self, device, lr):
        n = ceil4(self.size)
        self.n = n
        self.params = torch.zeros(n, device=device)
        self.grads = torch.zeros(n, device=device)
        self.exp_avg = torch.zeros(n, device=device)
        self.exp_avg_sq = torch.zeros(n, device=device)
        self.state = torch.tensor([0.0, 0.0, 1.0, 1.0, lr, 0.0, 0.0, 0.0], dtype=torch.float64, device=device)

    def view(self, name, buf="params"):
        off, rows, cols, ld = self.items[name]
        t = getattr(self, buf)[off:off + rows * ld].view(rows, ld)
        return t

    def ptr(self, name, buf="params", col=0):
        off = self.items[name][0]
        return getattr(self, buf).data_ptr() + 4 * (off + col)

    def ld(self, name):
        return self.items[name][3]

    def set_lr(self, lr):
        self.state[4] = lr

    def step_count(self):
        return int(self.state[1].item())


class Lin:
    def __init__(self, group, key, K, N, act):
        self.g, self.key, self.K, self.N, self.act = group, key, K, N, act
        group.add(key + ".weight", N, K)
        group.add(key + ".bias", 1, N)

    @property
    def ldw(self):
        return self.g.ld(self.key + ".weight")

    def w(self, buf="params", col=0):
        return self.g.ptr(self.key + ".weight", buf, col)

    def b(self, buf="params"):
        return self.g.ptr(self.key + ".bias", buf)


def _mlp(group, prefix, dims, final_act=0):
    """nn.Sequential(Linear, ELU, ..., Linear[, ELU]) with the reference's `<prefix>.<2i>` keys."""
    layers = []
    for i in range(len(dims) - 1):
        last = i == len(dims) - 2
        layers.append(Lin(group, f"{prefix}.{2 * i}", dims[i], dims[i + 1], final_act if last else 1))
    return layers


class Workspace:
    """Activation / gradient scratch for one batch size."""

    def __init__(self, device):
        self.device, self.t = device, {}

    def get(self, name, rows, cols):
        key = (name, rows, cols)
        if key not in self.t:
            self.t[key] = torch.zeros(rows, cols, device=self.device)
        return self.t[key]

    def ptr(self, name, rows, cols, col=0):
        return self.get(name, rows, cols).data_ptr() + 4 * col


def _p(t, col=0):
    return t.data_ptr() + 4 * col


class Kernels:
    """Thin typed wrappers over the C ABI (addresses are plain ints).

    precise=False (production): the tcgen05 / TMEM / TMA kernels (kind::tf32) wherever the shape allows
    (N >= 8, K >= 8), the mma.sync TF32 kernels otherwise.  precise=True (parity runs against the fp32 oracle):
    the 3xTF32 mma.sync kernels everywhere."""

    def __init__(self, precise=False, use_tc=True):
        self.lib = _lib.lib()
        self.precise = int(precise)
        self.use_tc = bool(use_tc) and not precise
        self.use_chain = False      # one launch per Linear / ELU chain at rollout sizes (measured slower: networks.CHAIN_MAX_ROWS)

    def fwd(self, lin, X, ldx, Y, ldy, M, act=None):
        act = lin.act if act is None else act
        if self.use_tc and lin.N >= 8 and lin.K >= 8:
            _lib.check(self.lib.b200_tc_linear_forward(X, ldx, lin.w(), lin.ldw, lin.b(), Y, ldy, M, lin.N, lin.K, act, _lib.stream_ptr()))
        else:
            _lib.check(self.lib.b200_linear_forward(X, ldx, lin.w(), lin.ldw, lin.b(), Y, ldy, M, lin.N, lin.K, act, self.precise,
                                                    _lib.stream_ptr()))

    def dgrad(self, lin, dY, lddy, Yprev, ldyp, dX, lddx, M, accumulate=0, wcol=0, K=None, dbias_prev=None):
        """dX[M,K] (+)= dY[M,N] . W[:, wcol:wcol+K] * elu'(Yprev).  With `dbias_prev` (the bias-gradient pointer of the layer
        below) the tcgen05 kernel also reduces the column sums of dX into it; returns True when it did."""
        K = lin.K if K is None else K
        if self.use_tc and lin.N >= 8 and K >= 8:
            if dbias_prev is not None and not accumulate:
                _lib.check(self.lib.b200_tc_linear_dgrad_bias(dY, lddy, lin.w(col=wcol), lin.ldw, Yprev, ldyp, dX, lddx, M, lin.N, K, 0,
                                                              dbias_prev, _lib.stream_ptr()))
                return True
            _lib.check(self.lib.b200_tc_linear_dgrad(dY, lddy, lin.w(col=wcol), lin.ldw, Yprev, ldyp, dX, lddx, M, lin.N, K, accumulate,
                                                     _lib.stream_ptr()))
        else:
            _lib.check(self.lib.b200_linear_dgrad(dY, lddy, lin.w(col=wcol), lin.ldw, Yprev, ldyp, dX, lddx, M, lin.N, K, accumulate,
                                                  self.precise, _lib.stream_ptr()))

    def wgrad(self, lin, dY, lddy, X, ldx, M, K=None, bias_done=False):
        """dW += dY^T . X and db += colsum(dY); `bias_done`: the dgrad of the layer above has already reduced db"""
        K = lin.K if K is None else K
        if self.use_tc and lin.N >= 8 and K >= 8 and M >= 32:
            _lib.check(self.lib.b200_tc_linear_wgrad(dY, lddy, X, ldx, lin.w("grads"), lin.ldw, M, lin.N, K, _lib.stream_ptr()))
            if not bias_done:
                _lib.check(self.lib.b200_colsum(dY, lddy, lin.b("grads"), M, lin.N, _lib.stream_ptr()))
        elif bias_done:
            raise RuntimeError("bias gradient was fused into the dgrad above but this layer's wgrad takes the fallback path")
        else:
            _lib.check(self.lib.b200_linear_wgrad(dY, lddy, X, ldx, lin.w("grads"), lin.ldw, lin.b("grads"), M, lin.N, K, self.precise,
                                                  _lib.stream_ptr()))


# One persistent launch for a whole chain (b200_tc_mlp_forward) is OFF: measured on B200 at M = 4096, replayed from CUDA graphs
# (tools/probe_chain.py): estimator chain 22.8 us as three launches vs 25.9 us as one, actor 34.6 vs 42.7, scan 14.6 vs 19.0 --
# a kernel boundary inside a graph costs ~2 us, the grid-wide barrier of the chain kernel (atomic + spin through L2) as much,
# and its one narrow tile shape for all layers is less efficient.  Kernels.use_chain = True switches it on (A/B, tests).
CHAIN_MAX_ROWS = 8192


def chain_forward(k, layers, ws, tag, X, ldx, out, ldo, M, max_ctas=74, after=None):
    """run a Linear/ELU chain; hidden activations live in ws under `<tag><i>`; the last layer writes (out, ldo).
    With `k.use_chain` (off by default, see CHAIN_MAX_ROWS) rollout-sized batches take ONE persistent launch for the whole
    chain -- the CTAs stay resident across the layers and meet at grid-wide barriers -- instead of one GEMM launch per layer;
    `max_ctas` is this chain's share of the 2 x SMs resident CTAs that concurrently running chains must fit in."""
    if (k.use_tc and getattr(k, "use_chain", False) and M <= CHAIN_MAX_ROWS and len(layers) <= 8 and ldx % 4 == 0 and X % 16 == 0
            and all(lin.K >= 8 for lin in layers)):
        arr = (_lib.MlpLayer * len(layers))()
        for i, lin in enumerate(layers):
            last = i == len(layers) - 1
            ldy = ldo if last else ceil4(lin.N)
            Y = out if last else ws.ptr(f"{tag}{i}", M, ldy)
            arr[i].W, arr[i].bias, arr[i].Y, arr[i].ldw, arr[i].ldy = lin.w(), lin.b(), Y, lin.ldw, ldy
            arr[i].N, arr[i].K, arr[i].act = lin.N, lin.K, lin.act
        sync = ws.ptr(f"chain_sync_{tag}", 1, 4)
        _lib.check(k.lib.b200_tc_mlp_forward(arr, len(layers), X, ldx, M, sync, int(max_ctas), _lib.stream_ptr()))
        if after is not None:                          # one launch: the hook can only run behind the whole chain
            after[1]()
        return
    for i, lin in enumerate(layers):
        last = i == len(layers) - 1
        if last:
            Y, ldy = out, ldo
        else:
            ldy = ceil4(lin.N)
            Y = ws.ptr(f"{tag}{i}", M, ldy)
        k.fwd(lin, X, ldx, Y, ldy, M)
        X, ldx = Y, ldy
        if after is not None and after[0] == i:       # (layer index, callback): e.g. fork another stream behind this layer
            after[1]()


def chain_backward(k, layers, ws, tag, X0, ldx0, dOut, lddo, M, need_dx0=False, dX0=None, lddx0=0, wgrad_on=None):
    """backward of chain_forward: wgrad of every layer, dgrad between layers (elu' fused from the stored activation).
    `wgrad_on(fn)`: run the weight-gradient launches `fn` elsewhere (another stream, ordered after everything queued here so
    far): a layer's wgrad and dgrad both only READ dY, so the dgrad chain -- the critical path -- need not wait for the wgrads."""
    dY, lddy = dOut, lddo
    bias_done = False
    for i in reversed(range(len(layers))):
        lin = layers[i]
        if i > 0:
            ldx = ceil4(layers[i - 1].N)
            X = ws.ptr(f"{tag}{i - 1}", M, ldx)
        else:
            X, ldx = X0, ldx0
        if wgrad_on is not None:
            wgrad_on(lambda lin=lin, dY=dY, lddy=lddy, X=X, ldx=ldx, bd=bias_done: k.wgrad(lin, dY, lddy, X, ldx, M, bias_done=bd))
        else:
            k.wgrad(lin, dY, lddy, X, ldx, M, bias_done=bias_done)
        bias_done = False
        if i > 0:
            dX = ws.ptr(f"d{tag}{i - 1}", M, ldx)
            below = layers[i - 1]
            fuse = k.use_tc and below.N >= 8 and below.K >= 8 and M >= 32      # the layer below takes the tcgen05 wgrad path
            bias_done = bool(k.dgrad(lin, dY, lddy, X, ldx, dX, ldx, M, dbias_prev=below.b("grads") if fuse else None))
            dY, lddy = dX, ldx
        elif need_dx0:
            k.dgrad(lin, dY, lddy, None, 0, dX0, lddx0, M)


def _uniform(shape, bound, gen):
    return (torch.rand(shape, generator=gen) * 2 - 1) * bound


class ActorCritic:
    is_recurrent = False

    def __init__(self, num_proprio, num_privileged_obs, num_critic_obs, num_estimated_obs, num_scan_obs, num_actions,
                 history_buffer_length, actor_hidden_dims=[256, 256, 256], critic_hidden_dims=[256, 256, 256],
                 priv_encoder_hidden_dims=[64, 20], scan_encoder_hidden_dims=[128, 64], latent_encoder_output_dim=20,
                 scan_encoder_output_dim=32, activation='elu', init_noise_std=1.0, device="cuda:0", seed=0, precise=False,
                 learning_rate=1e-3, **kwargs):
        if activation != 'elu':
            raise NotImplementedError("the fused kernels implement ELU (the only activation the go2 configs use)")
        if num_proprio != 52 or history_buffer_length != 10:
            raise NotImplementedError("AdaptationEncoder geometry is the reference's: 10 steps x 52 proprio")
        self.num_proprio, self.num_privileged_obs, self.num_critic_obs = num_proprio, num_privileged_obs, num_critic_obs
        self.num_estimated_obs, self.num_scan_obs, self.num_actions = num_estimated_obs, num_scan_obs, num_actions
        self.history_buffer_length = history_buffer_length
        self.latent_dim, self.scan_latent_dim = latent_encoder_output_dim, scan_encoder_output_dim
        self.device = torch.device(device)
        self.k = Kernels(precise)
        self.num_obs = num_proprio * (history_buffer_length + 1)
        # actor input = [obs | latent | scan latent | estimated obs] (actor_critic.py:79, :195)
        self.col_latent = self.num_obs
        self.col_scan = self.col_latent + latent_encoder_output_dim
        self.col_est = self.col_scan + scan_encoder_output_dim
        self.actor_in_dim = self.col_est + num_estimated_obs
        self.ld_actor_in = ceil4(self.actor_in_dim)
        assert self.col_latent % 4 == 0 and self.col_scan % 4 == 0 and self.col_est % 4 == 0

        g = self.main = FlatGroup()
        self.actor = _mlp(g, "actor", [self.actor_in_dim] + list(actor_hidden_dims) + [num_actions])
        self.critic = _mlp(g, "critic", [num_critic_obs] + list(critic_hidden_dims) + [1])
        self.priv = _mlp(g, "privileged_encoder_.priv_encoder", [num_privileged_obs] + list(priv_encoder_hidden_dims) + [latent_encoder_output_dim])
        g.add("std", 1, num_actions)
        self.scan = _mlp(g, "scan_encoder.scan_encoder", [num_scan_obs] + list(scan_encoder_hidden_dims) + [scan_encoder_output_dim])
        g.finalize(self.device, learning_rate)
        a = self.adapt = FlatGroup()
        self.ad_fc = Lin(a, "adaptation_encoder_.fc_encoder.0", num_proprio, 30, 1)
        self.ad_c1 = Lin(a, "adaptation_encoder_.conv_layers.0", 4 * 32, 20, 1)
        self.ad_c2 = Lin(a, "adaptation_encoder_.conv_layers.2", 2 * 20, 10, 1)
        self.ad_out = Lin(a, "adaptation_encoder_.fc_final.0", 3 * 12, latent_encoder_output_dim, 1)
        a.finalize(self.device, learning_rate)
        self._ws = {}
        self._last = None
        self.fused_adapt = True
        self.load_state_dict(self._random_state_dict(init_noise_std, seed))

    # ---- parameters ---------------------------------------------------------------------------------
    @property
    def std(self):
        return self.main.view("std")[0]

    def _ref_shapes(self):
        """checkpoint key -> shape, in the reference's parameter order."""
        shapes = OrderedDict()
        shapes["std"] = (self.num_actions,)
        for lin in self.actor + self.critic:
            shapes[lin.key + ".weight"], shapes[lin.key + ".bias"] = (lin.N, lin.K), (lin.N,)
        shapes["adaptation_encoder_.fc_encoder.0.weight"], shapes["adaptation_encoder_.fc_encoder.0.bias"] = (30, 52), (30,)
        shapes["adaptation_encoder_.conv_layers.0.weight"], shapes["adaptation_encoder_.conv_layers.0.bias"] = (20, 30, 4), (20,)
        shapes["adaptation_encoder_.conv_layers.2.weight"], shapes["adaptation_encoder_.conv_layers.2.bias"] = (10, 20, 2), (10,)
        shapes["adaptation_encoder_.fc_final.0.weight"], shapes["adaptation_encoder_.fc_final.0.bias"] = (self.latent_dim, 30), (self.latent_dim,)
        for lin in self.priv + self.scan:
            shapes[lin.key + ".weight"], shapes[lin.key + ".bias"] = (lin.N, lin.K), (lin.N,)
        return shapes

    def _random_state_dict(self, init_noise_std, seed):
        gen = torch.Generator().manual_seed(seed)
        sd = OrderedDict()
        for key, shape in self._ref_shapes().items():
            if key == "std":
                sd[key] = init_noise_std * torch.ones(shape)
                continue
            if key.endswith(".weight"):
                fan_in = int(torch.tensor(shape[1:]).prod())
                self._fan = fan_in
            sd[key] = _uniform(shape, 1.0 / math.sqrt(self._fan), gen)      # nn.Linear / nn.Conv1d default init
        return sd

    def _group_of(self, key):
        return self.adapt if key.startswith("adaptation_encoder_") else self.main

    def state_dict(self):
        sd = OrderedDict()
        for key, shape in self._ref_shapes().items():
            g = self._group_of(key)
            v = g.view(key).detach()
            if key == "std" or key.endswith(".bias"):
                sd[key] = v[0, :shape[0]].clone()
            elif key == "adaptation_encoder_.conv_layers.0.weight":
                sd[key] = v[:, :128].view(20, 4, 32)[:, :, :30].permute(0, 2, 1).contiguous()
            elif key == "adaptation_encoder_.conv_layers.2.weight":
                sd[key] = v[:, :40].view(10, 2, 20).permute(0, 2, 1).contiguous()
            elif key == "adaptation_encoder_.fc_final.0.weight":
                sd[key] = v[:, :36].view(self.latent_dim, 3, 12)[:, :, :10].permute(0, 2, 1).reshape(self.latent_dim, 30).contiguous()
            else:
                sd[key] = v[:, :shape[1]].clone()
        return sd

    def load_state_dict(self, sd, strict=True):
        shapes = self._ref_shapes()
        missing = [k for k in shapes if k not in sd]
        if strict and missing:
            raise KeyError(f"missing keys in state_dict: {missing}")
        for key, shape in shapes.items():
            if key not in sd:
                continue
            src = torch.as_tensor(sd[key]).detach().to(self.device, torch.float32)
            assert tuple(src.shape) == tuple(shape), (key, tuple(src.shape), shape)
            v = self._group_of(key).view(key)
            v.zero_()
            if key == "std" or key.endswith(".bias"):
                v[0, :shape[0]] = src
            elif key == "adaptation_encoder_.conv_layers.0.weight":
                v[:, :128].view(20, 4, 32)[:, :, :30] = src.permute(0, 2, 1)
            elif key == "adaptation_encoder_.conv_layers.2.weight":
                v[:, :40].view(10, 2, 20)[:] = src.permute(0, 2, 1)
            elif key == "adaptation_encoder_.fc_final.0.weight":
                v[:, :36].view(self.latent_dim, 3, 12)[:, :, :10] = src.view(self.latent_dim, 10, 3).permute(0, 2, 1)
            else:
                v[:, :shape[1]] = src

    def parameters(self):
        return [self.main.params, self.adapt.params]

    def train(self):
        return self

    def eval(self):
        return self

    def test(self):
        return self

    def to(self, device):
        assert torch.device(device).type == "cuda"
        return self

    def reset(self, dones=None):
        pass

    # ---- sub-network forward passes on raw addresses ------------------------------------------------
    def ws(self, M):
        if M not in self._ws:
            self._ws[M] = Workspace(self.device)
        return self._ws[M]

    # Budgets of the one-launch chains (CTAs; an SM holds two, 296 on a B200): the estimator (100), the two encoders (48 each)
    # and the critic (74) may run concurrently, then the actor (148) beside the critic -- never more than 270 in flight.
    def fwd_priv(self, ws, X, ldx, out, ldo, M):
        chain_forward(self.k, self.priv, ws, "p", X, ldx, out, ldo, M, max_ctas=48)

    def fwd_scan(self, ws, X, ldx, out, ldo, M):
        chain_forward(self.k, self.scan, ws, "s", X, ldx, out, ldo, M, max_ctas=48)

    def fwd_actor(self, ws, X, ldx, out, ldo, M, after=None):
        chain_forward(self.k, self.actor, ws, "a", X, ldx, out, ldo, M, max_ctas=148, after=after)

    def fwd_critic(self, ws, X, ldx, out, ldo, M):
        chain_forward(self.k, self.critic, ws, "c", X, ldx, out, ldo, M, max_ctas=74)

    def fwd_adapt(self, ws, X, ldx, out, ldo, M, save=False):
        """AdaptationEncoder.forward (support_networks.py:128-175) on obs rows (history = first 520 columns): one fused
        fp32 kernel; `save` also stores the hidden activations for bwd_adapt (DAgger)."""
        if self.fused_adapt:
            a, f, c1_, c2_, o = self.adapt, self.ad_fc, self.ad_c1, self.ad_c2, self.ad_out
            sv = (ws.ptr("ad_proj", M, 320), ws.ptr("ad_c1", M, 80), ws.ptr("ad_c2", M, 36)) if save else (None, None, None)
            _lib.check(self.k.lib.b200_adaptation_forward(X, ldx, f.w(), f.b(), c1_.w(), c1_.b(), c2_.w(), c2_.b(), o.w(), o.b(), out, ldo,
                                                          sv[0], sv[1], sv[2], M, _lib.stream_ptr()))
            return
        self._fwd_adapt_gemms(ws, X, ldx, out, ldo, M)

    def _fwd_adapt_gemms(self, ws, X, ldx, out, ldo, M):
        """the same network as 18 Linear-kernel launches (kept as the cross-check of the fused kernel)."""
        k, NP = self.k, self.num_proprio
        proj = ws.ptr("ad_proj", M, 320)
        for t in range(10):                                    # fc_encoder on each history step
            k.fwd(self.ad_fc, X + 4 * NP * t, ldx, proj + 4 * 32 * t, 320, M)
        c1 = ws.ptr("ad_c1", M, 80)
        for t in range(4):                                     # Conv1d(30, 20, k=4, s=2)
            k.fwd(self.ad_c1, proj + 4 * 64 * t, 320, c1 + 4 * 20 * t, 80, M)
        c2 = ws.ptr("ad_c2", M, 36)
        for t in range(3):                                     # Conv1d(20, 10, k=2, s=1)
            k.fwd(self.ad_c2, c1 + 4 * 20 * t, 80, c2 + 4 * 12 * t, 36, M)
        k.fwd(self.ad_out, c2, 36, out, ldo, M)                # Flatten + fc_final

    def bwd_adapt(self, ws, X, ldx, dOut, lddo, out, ldo, M):
        """backward of fwd_adapt into self.adapt.grads; dOut is d(loss)/d(latent) BEFORE the final ELU'."""
        k, lib, st = self.k, self.k.lib, _lib.stream_ptr
        _lib.check(lib.b200_elu_backward(dOut, lddo, out, ldo, M, self.latent_dim, st()))
        proj, c1, c2 = ws.ptr("ad_proj", M, 320), ws.ptr("ad_c1", M, 80), ws.ptr("ad_c2", M, 36)
        dproj, dc1, dc2 = ws.get("d_ad_proj", M, 320), ws.get("d_ad_c1", M, 80), ws.get("d_ad_c2", M, 36)
        k.wgrad(self.ad_out, dOut, lddo, c2, 36, M)
        k.dgrad(self.ad_out, dOut, lddo, c2, 36, _p(dc2), 36, M)
        dc1.zero_()
        for t in range(3):
            k.wgrad(self.ad_c2, _p(dc2, 12 * t), 36, c1 + 4 * 20 * t, 80, M)
            k.dgrad(self.ad_c2, _p(dc2, 12 * t), 36, c1 + 4 * 20 * t, 80, _p(dc1, 20 * t), 80, M, accumulate=1)
        dproj.zero_()
        for t in range(4):
            k.wgrad(self.ad_c1, _p(dc1, 20 * t), 80, proj + 4 * 64 * t, 320, M)
            k.dgrad(self.ad_c1, _p(dc1, 20 * t), 80, proj + 4 * 64 * t, 320, _p(dproj, 64 * t), 320, M, accumulate=1)
        for t in range(10):
            k.wgrad(self.ad_fc, _p(dproj, 32 * t), 320, X + 4 * self.num_proprio * t, ldx, M)

    # ---- the reference's public methods (tensor in / tensor out) ------------------------------------
    def _pack_actor_in(self, ws, obs, M):
        buf = ws.get("actor_in", M, self.ld_actor_in)
        buf[:, :self.num_obs].copy_(obs)
        return buf

    def privileged_encoder(self, privileged_obs_buf):
        M = privileged_obs_buf.shape[0]
        ws = self.ws(M)
        x = ws.get("priv_in", M, ceil4(self.num_privileged_obs))
        x[:, :self.num_privileged_obs].copy_(privileged_obs_buf)
        out = ws.get("latent_out", M, self.latent_dim)
        self.fwd_priv(ws, _p(x), x.shape[1], _p(out), self.latent_dim, M)
        return out

    def adaptation_encoder(self, obs_buf):
        M = obs_buf.shape[0]
        ws = self.ws(M)
        x = self._pack_actor_in(ws, obs_buf, M)
        out = ws.get("latent_out", M, self.latent_dim)
        self.fwd_adapt(ws, _p(x), self.ld_actor_in, _p(out), self.latent_dim, M)
        return out

    def get_latent(self, obs_buf, privileged_obs_buf, adaptation_mode=False):
        return self.adaptation_encoder(obs_buf) if adaptation_mode else self.privileged_encoder(privileged_obs_buf)

    def _actor_mean(self, obs_buf, privileged_obs_buf, estimated_obs_buf, scan_obs_buf, adaptation_mode):
        M = obs_buf.shape[0]
        ws = self.ws(M)
        x = self._pack_actor_in(ws, obs_buf, M)
        ld = self.ld_actor_in
        if adaptation_mode:
            self.fwd_adapt(ws, _p(x), ld, _p(x, self.col_latent), ld, M)
        else:
            pin = ws.get("priv_in", M, ceil4(self.num_privileged_obs))
            pin[:, :self.num_privileged_obs].copy_(privileged_obs_buf)
            self.fwd_priv(ws, _p(pin), pin.shape[1], _p(x, self.col_latent), ld, M)
        sin = scan_obs_buf.contiguous()
        self.fwd_scan(ws, _p(sin), sin.shape[1], _p(x, self.col_scan), ld, M)
        x[:, self.col_est:self.col_est + self.num_estimated_obs].copy_(estimated_obs_buf)
        mu = ws.get("mu", M, self.num_actions)
        self.fwd_actor(ws, _p(x), ld, _p(mu), self.num_actions, M)
        return mu

    def update_distribution(self, obs_buf, privileged_obs_buf, estimated_obs_buf, scan_obs_buf, adaptation_mode=False):
        mu = self._actor_mean(obs_buf, privileged_obs_buf, estimated_obs_buf, scan_obs_buf, adaptation_mode)
        self._last = dict(mu=mu, sigma=self.std.unsqueeze(0).expand_as(mu))

    def act(self, obs_buf, privileged_obs_buf, estimated_obs_buf, scan_obs_buf, adaptation_mode=False, seed=0, step=0):
        self.update_distribution(obs_buf, privileged_obs_buf, estimated_obs_buf, scan_obs_buf, adaptation_mode)
        mu = self._last["mu"]
        M, A = mu.shape
        ws = self.ws(M)
        actions, logp = ws.get("actions", M, A), ws.get("logp", M, 1)
        _lib.check(self.k.lib.b200_sample_actions(_p(mu), A, self.main.ptr("std"), seed, step, _p(actions), _p(logp), None, None,
                                                  M, A, _lib.stream_ptr()))
        self._last["actions"], self._last["logp"] = actions, logp[:, 0]
        return actions

    def act_inference(self, obs_buf, privileged_obs_buf, estimated_obs_buf, scan_obs_buf, adaptation_mode=False):
        return self._actor_mean(obs_buf, privileged_obs_buf, estimated_obs_buf, scan_obs_buf, adaptation_mode)

    @property
    def action_mean(self):
        return self._last["mu"]

    @property
    def action_std(self):
        return self._last["sigma"]

    @property
    def entropy(self):
        s = self._last["sigma"]
        return (0.5 + 0.5 * math.log(2 * math.pi) + torch.log(s)).sum(dim=-1)

    def get_actions_log_prob(self, actions):
        mu, s = self._last["mu"], self._last["sigma"]
        return (-((actions - mu) ** 2) / (2 * s ** 2) - torch.log(s) - math.log(math.sqrt(2 * math.pi))).sum(dim=-1)

    def evaluate(self, critic_observations, **kwargs):
        M = critic_observations.shape[0]
        ws = self.ws(M)
        x = critic_observations.contiguous()
        out = ws.get("value", M, 1)
        self.fwd_critic(ws, _p(x), x.shape[1], _p(out), 1, M)
        return out


class MlpEstimator:
    def __init__(self, num_proprio, history_buffer_length, output_dim, hidden_dims=[128, 64], activation="elu", use_history=True,
                 device="cuda:0", seed=1, precise=False, learning_rate=1e-3):
        if activation != "elu":
            raise NotImplementedError("ELU only")
        self.use_history, self.num_proprio, self.history_buffer_length = use_history, num_proprio, history_buffer_length
        self.input_dim = num_proprio * (history_buffer_length + 1) if use_history else num_proprio
        self.in_col = 0 if use_history else num_proprio * history_buffer_length
        self.output_dim = output_dim
        self.device = torch.device(device)
        self.k = Kernels(precise)
        g = self.group = FlatGroup()
        self.layers = _mlp(g, "estimator", [self.input_dim] + list(hidden_dims) + [output_dim])
        g.finalize(self.device, learning_rate)
        self._ws = {}
        gen = torch.Generator().manual_seed(seed)
        sd = OrderedDict()
        for lin in self.layers:
            b = 1.0 / math.sqrt(lin.K)
            sd[lin.key + ".weight"], sd[lin.key + ".bias"] = _uniform((lin.N, lin.K), b, gen), _uniform((lin.N,), b, gen)
        self.load_state_dict(sd)

    def state_dict(self):
        sd = OrderedDict()
        for lin in self.layers:
            sd[lin.key + ".weight"] = self.group.view(lin.key + ".weight")[:, :lin.K].clone()
            sd[lin.key + ".bias"] = self.group.view(lin.key + ".bias")[0].clone()
        return sd

    def load_state_dict(self, sd, strict=True):
        for lin in self.layers:
            w = self.group.view(lin.key + ".weight")
            w.zero_()
            w[:, :lin.K] = torch.as_tensor(sd[lin.key + ".weight"]).to(self.device, torch.float32)
            self.group.view(lin.key + ".bias")[0] = torch.as_tensor(sd[lin.key + ".bias"]).to(self.device, torch.float32)

    def parameters(self):
        return [self.group.params]

    def to(self, device):
        return self

    def ws(self, M):
        if M not in self._ws:
            self._ws[M] = Workspace(self.device)
        return self._ws[M]

    def fwd(self, ws, X, ldx, out, ldo, M):
        chain_forward(self.k, self.layers, ws, "e", X + 4 * self.in_col, ldx, out, ldo, M, max_ctas=100)

    def forward(self, obs_with_history):
        M = obs_with_history.shape[0]
        ws = self.ws(M)
        x = obs_with_history
        if x.shape[1] % 4 or not x.is_contiguous():
            buf = ws.get("in", M, ceil4(x.shape[1]))
            buf[:, :x.shape[1]].copy_(x)
            x = buf
        out = ws.get("out", M, self.output_dim)
        self.fwd(ws, _p(x), x.shape[1], _p(out), self.output_dim, M)
        return out

    __call__ = forward
