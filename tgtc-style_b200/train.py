"""Data-parallel training step of the NeRF pair, host side.

Replaces the body of Origin_train's loop (train_tgtcs.py:222-276) for perturb=0 / sigma_noise_std=0:
    forward coarse+fine -> loss = mse(coarse)+mse(fine) -> backward -> Adam step -> lr decay
The forward/backward arithmetic is one C-ABI call (tgtc_train_step); this module owns only the plumbing the
reference leaves to torch: the fp32 master parameters (nn.Linear layout, the reference's own state_dict names), the
optimizer (Adam as train_tgtcs.py:39 -- one fused kernel over the flat buffers, tgtc_adam_step) and -- with more than one rank -- the single gradient
all-reduce (NCCL sum over the flat 2 x 595 844-float buffer; the 1/N of the mean is already folded into the loss
through n_total).  Rays shard across ranks with dist.shard_range; nothing else is exchanged.
"""
import torch
import torch.distributed as dist

from .dist import shard_range
from .render import LAYER_NAMES, LAYER_SHAPES


class NerfTrainer:
    def __init__(self, renderer, coarse, fine, lr=5e-4, lr_decay_steps=100000, lr_decay_rate=0.1, group=None, max_rays_per_pass=8192):
        """coarse / fine: state_dicts (or nn.Modules) with the reference's parameter names (models.py:75-91).
        lr schedule: lr * rate^(step/decay_steps) as train_tgtcs.py:272-276."""
        self.r = renderer
        self.group = group
        self.max_rays = int(max_rays_per_pass)
        dev = renderer.device
        # fp32 masters of both nets live in ONE flat buffer laid out like the gradient buffer; the per-parameter tensors are
        # views into it (state_dict export, re-packing) and the optimizer is one fused kernel over the flat buffers
        self.flat = torch.zeros(2 * 595844, dtype=torch.float32, device=dev)
        self.grads = torch.zeros_like(self.flat)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        pviews = renderer.grad_views(self.flat)
        self.params = []
        for src, views in zip((coarse, fine), pviews):
            sd = src.state_dict() if hasattr(src, "state_dict") else src
            for name in LAYER_NAMES:
                for suffix in (".weight", ".bias"):
                    views[name + suffix].copy_(sd[name + suffix].detach().to(dev, torch.float32))
            self.params.append(views)
        self.fused = hasattr(renderer, "adam_step")
        if not self.fused:                      # host-logic tests with a stand-in renderer: torch's own Adam on the same views
            gviews = renderer.grad_views(self.grads)
            plist = []
            for d, gv in zip(self.params, gviews):
                for k in d:
                    d[k] = torch.nn.Parameter(d[k])
                    d[k].grad = gv[k]
                    plist.append(d[k])
            self.opt = torch.optim.Adam(plist, lr=lr, betas=(0.9, 0.999))
        self.lr0, self.decay_steps, self.decay_rate = lr, lr_decay_steps, lr_decay_rate
        self.lr = lr
        self.step_count = 0
        self.r.set_weights(self.params[0], self.params[1])

    def world(self):
        return dist.get_world_size(self.group) if dist.is_initialized() else 1

    def rank(self):
        return dist.get_rank(self.group) if dist.is_initialized() else 0

    def forward_backward(self, rays_o, rays_d, rgb_gt, n_total=None, perturb=False, sigma_noise_std=0.):
        """Gradients of this rank's rays into self.grads (ray chunks of max_rays_per_pass accumulate).  Returns the
        local loss contribution (device scalar; summing it over ranks gives the step's loss)."""
        n = rays_o.shape[0]
        n_total = n if n_total is None else n_total
        loss = None
        for b in range(0, max(n, 1), self.max_rays):
            e = min(n, b + self.max_rays)
            rand = nzc = nzf = None
            if perturb:             # the generator call of utils.py:519-520
                rand = torch.zeros([e - b, 64], device=rays_o.device)
                torch.nn.init.uniform_(rand, 0, 1)
            if sigma_noise_std > 0:  # the generator calls of utils.py:373-374 (coarse pass, then fine pass)
                nzc = torch.randn([e - b, 64], device=rays_o.device) * sigma_noise_std
                nzf = torch.randn([e - b, 128], device=rays_o.device) * sigma_noise_std
            out = self.r.train_step(rays_o[b:e], rays_d[b:e], rgb_gt[b:e], n_total=n_total, grads=self.grads, accumulate=b > 0,
                                    rand=rand, noise_coarse=nzc, noise_fine=nzf)
            loss = out["loss"] if loss is None else loss + out["loss"]
        return loss

    def step(self, rays_o, rays_d, rgb_gt, sharded=False, perturb=False, sigma_noise_std=0.):
        """One optimisation step.  rays: this step's global batch (every rank passes the same tensors and takes its
        shard_range) or, with sharded=True, this rank's own shard of a global batch of world * n rays."""
        world = self.world()
        if sharded or world == 1:
            ro, rd, gt = rays_o, rays_d, rgb_gt
            n_total = rays_o.shape[0] * world
        else:
            b, e = shard_range(rays_o.shape[0], self.rank(), world)
            ro, rd, gt = rays_o[b:e], rays_d[b:e], rgb_gt[b:e]
            n_total = rays_o.shape[0]
        loss = self.forward_backward(ro, rd, gt, n_total, perturb=perturb, sigma_noise_std=sigma_noise_std)
        if world > 1:
            dist.all_reduce(self.grads, op=dist.ReduceOp.SUM, group=self.group)   # the path's only collective
        self.step_count += 1
        if self.fused:
            self.r.adam_step(self.flat, self.grads, self.exp_avg, self.exp_avg_sq, self.step_count, lr=self.lr)
        else:
            self.opt.step()
        self.lr = self.lr0 * (self.decay_rate ** (self.step_count / self.decay_steps))  # train_tgtcs.py:272-276
        if not self.fused:
            for g in self.opt.param_groups:
                g["lr"] = self.lr
        self.r.set_weights(self.params[0], self.params[1])    # re-pack the bf16 / transposed images from the fp32 masters
        return loss

    def state_dicts(self):
        return tuple({k: p.detach() for k, p in d.items()} for d in self.params)
