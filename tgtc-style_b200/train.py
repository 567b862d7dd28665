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
    def __init__(self, renderer, coarse, fine, lr=5e-4, lr_decay_steps=100000, lr_decay_rate=0.1, group=None, max_rays_per_pass=8192,
                 seed=None, overlap_allreduce=True):
        """coarse / fine: state_dicts (or nn.Modules) with the reference's parameter names (models.py:75-91).
        lr schedule: lr * rate^(step/decay_steps) as train_tgtcs.py:272-276."""
        self.r = renderer
        self.group = group
        self.max_rays = int(max_rays_per_pass)
        self.seed = seed   # int: perturb / sigma noise are drawn inside the kernels (Philox) instead of by torch's generator
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
        # data-parallel overlap: the coarse net's gradient half is all-reduced on a side stream while the fine
        # net's forward / backward still runs; the fine half follows when the step's work is done
        self.overlap = bool(overlap_allreduce) and hasattr(renderer, "set_coarse_event") and torch.cuda.is_available()
        self._comm = None

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
            if self.seed is not None and (perturb or sigma_noise_std > 0):
                # one Philox key per (step, rank, ray chunk): elements are indexed from 0 inside every library call
                key = (int(self.seed) * 0x9E3779B97F4A7C15 + (self.step_count << 24) + (self.rank() << 12) + b // self.max_rays) & (2 ** 64 - 1)
                out = self.r.train_step(rays_o[b:e], rays_d[b:e], rgb_gt[b:e], n_total=n_total, grads=self.grads, accumulate=b > 0,
                                        seed=key, perturb=perturb, sigma_noise_std=sigma_noise_std)
                loss = out["loss"] if loss is None else loss + out["loss"]
                continue
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
        if world > 1 and self.overlap:
            loss = self._forward_backward_overlapped(ro, rd, gt, n_total, perturb, sigma_noise_std)
        else:
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

    def _forward_backward_overlapped(self, ro, rd, gt, n_total, perturb, sigma_noise_std):
        """forward_backward + the gradient all-reduce in two halves: the coarse net's half starts (on a side stream) as soon as
        the library signals that it is final, underneath the fine net's forward / backward; the fine half follows the step."""
        dev = self.r.device
        cur = torch.cuda.current_stream(dev)
        if self._comm is None:
            self._comm = torch.cuda.Stream(device=dev)
            self._ev_coarse = torch.cuda.Event()
            self._ev_end = torch.cuda.Event()
        P = self.grads.numel() // 2
        self.r.set_coarse_event(self._ev_coarse)
        loss = self.forward_backward(ro, rd, gt, n_total, perturb=perturb, sigma_noise_std=sigma_noise_std)
        self._ev_end.record(cur)
        self._comm.wait_event(self._ev_coarse)          # the last ray chunk's record: every chunk's coarse gradient is in
        with torch.cuda.stream(self._comm):
            dist.all_reduce(self.grads[:P], op=dist.ReduceOp.SUM, group=self.group)
            self._comm.wait_event(self._ev_end)
            dist.all_reduce(self.grads[P:], op=dist.ReduceOp.SUM, group=self.group)
        cur.wait_stream(self._comm)                      # the optimizer step reads the reduced gradients
        return loss

    def state_dicts(self):
        return tuple({k: p.detach() for k, p in d.items()} for d in self.params)


# ----------------------------------------------------------------------------------------------------------------------
# Style_train (train_tgtcs.py:311-495): second training phase -- the two style modules and the per-(style, frame) latents
# learn on frozen NeRF nets.

def cosine_similarity_rows(a, b):
    """VGGNet.cosine_similarity (VGGNet.py:204-210): per-row cosine of two [N,3] maps."""
    an = a / (torch.norm(a, dim=1, keepdim=True) + 1e-8)
    bn = b / (torch.norm(b, dim=1, keepdim=True) + 1e-8)
    return torch.sum(an * bn, dim=1)


def l2_norm(x):
    """utils.L2_norm (utils.py:459)."""
    return torch.sqrt(torch.sum(x ** 2) + 1e-8)


class StyleLatents:
    """Host mirror of models.StyleLatents_variational (models.py:475-549): a [style_num, frame_num, 32] table of latents, per-style
    mu / logvar, `forward` = mu + sigma_scale * (latents[id] - mu), `minus_logp`, Adam(lr=1e-3) on the table.  Tiny tensors:
    plain torch on the device (plumbing); the per-ray latents it returns feed tgtc_style_train_forward."""

    def __init__(self, latents, mu, logvar, dataset_type="llff", sigma_scale=1.0, lr=1e-3, renderer=None):
        self.latents = latents.detach().clone().contiguous().requires_grad_(True)
        self.mu, self.logvar = mu.detach(), logvar.detach()
        self.frame_num = latents.shape[1]
        self.dataset_type = dataset_type
        self.sigma_scale = sigma_scale
        self.lr = lr
        # Adam(lr=1e-3) on the table (models.py:541-542): the library's fused kernel when a renderer is given, else torch's
        self.r = renderer
        self.step_count = 0
        if renderer is not None:
            self.exp_avg = torch.zeros_like(self.latents)
            self.exp_avg_sq = torch.zeros_like(self.latents)
            self.opt = None
        else:
            self.opt = torch.optim.Adam([self.latents], lr=lr)

    def __call__(self, style_ids, frame_ids):
        flat = style_ids * self.frame_num + frame_ids
        tab = self.latents.reshape(-1, self.latents.shape[-1])
        if self.dataset_type == "llff":
            flat = flat % tab.shape[0]      # == indexing the 7x tiled table of models.py:496, without the copy
        mu = self.mu[style_ids]
        return mu + self.sigma_scale * (tab[flat] - mu)              # models.py:506

    def minus_logp(self, style_ids, frame_ids, lat=None):
        if lat is None:
            lat = self(style_ids, frame_ids)
        mu, logvar = self.mu[style_ids], self.logvar[style_ids]
        return torch.sum((lat - mu) ** 2 / (torch.exp(0.5 * logvar) + 1e-3), -1).mean()   # models.py:531-537

    def use_fused_adam(self, renderer):
        if self.opt is not None and self.step_count == 0 and hasattr(renderer, "adam_step"):
            self.r = renderer
            self.exp_avg = torch.zeros_like(self.latents)
            self.exp_avg_sq = torch.zeros_like(self.latents)
            self.opt = None

    def fused_ok(self):
        return self.r is not None and hasattr(self.r, "style_latents_forward")

    def forward_fused(self, style_ids, frame_ids, n_logp):
        """-> (per-ray latents [n,32] without autograd history, minus_logp SUM of the first n_logp rays [1])"""
        return self.r.style_latents_forward(self.latents.detach(), self.mu, self.logvar, style_ids, frame_ids, n_logp, self.frame_num,
                                            self.sigma_scale, table_tiles=7 if self.dataset_type == "llff" else 1)

    def backward_fused(self, style_ids, frame_ids, n_logp, dlat, logp_scale):
        """table gradient of <dlat, lat> + logp_scale * logp_sum, written to self.latents.grad"""
        self.latents.grad = self.r.style_latents_backward(self.latents.detach(), self.mu, self.logvar, style_ids, frame_ids, n_logp,
                                                          self.frame_num, dlat, logp_scale, self.sigma_scale,
                                                          table_tiles=7 if self.dataset_type == "llff" else 1)

    def zero_grad(self):
        self.latents.grad = None

    def step(self):
        if self.latents.grad is None:
            return
        self.step_count += 1
        if self.opt is not None:
            self.opt.step()
        else:
            with torch.no_grad():
                self.r.adam_step(self.latents.view(-1), self.latents.grad.contiguous().view(-1), self.exp_avg.view(-1),
                                 self.exp_avg_sq.view(-1), self.step_count, lr=self.lr)


class StyleTrainer:
    """One Style_train iteration (train_tgtcs.py:354-495) per `step`:
         batch 2 (the loss_coh batch, unshuffled) and batch 1 (shuffled): forward of the frozen NeRF nets + both style modules
         with per-ray latents and perturbed samples (tgtc_style_train_forward), compositing, resampling, fine pass;
         loss = lambda_rgb (mse(coarse) + mse(fine)) + lambda_logp(step) * minus_logp  [+ lambda_coh * loss_coh];
         backward into the style modules (tgtc_style_train_backward) and the latents; Adam on both.
    With a real renderer every operation of the iteration is a library call (style_loss.cu for the losses on the [N,3] maps and
    the latent model); the same losses written with torch ops + autograd are the second path (`fused_losses=False`, stand-in
    renderers).  Two deviations from the reference loop, both forced by it not running as written on
    torch >= 1.5 (it backpropagates twice through a graph whose weights the first optimizer step modified in place, and
    through the previous iteration's graph via `x` / `y`): the previous batch's maps enter loss_coh as constants, and one
    backward serves both optimizers."""

    def __init__(self, renderer, concat_style, style, latents, lr=5e-4, rgb_loss_lambda=1.0, logp_loss_lambda=0.1, logp_loss_decay=1.0,
                 loss_coh_lambda=1e2, origin_step=0, frame_num=None, sigma_noise_std=0.0, group=None, seed=None, fused_losses=True):
        self.r = renderer
        self.group = group
        self.lat = latents
        self.lat.use_fused_adam(renderer)
        dev = renderer.device
        P = renderer.style_num_params()
        self.flat = torch.zeros(P, dtype=torch.float32, device=dev)
        self.grads = torch.zeros_like(self.flat)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.params = renderer.style_grad_views(self.flat)
        for src, views in zip((concat_style, style), self.params):
            sd = src.state_dict() if hasattr(src, "state_dict") else src
            for k in views:
                views[k].copy_(sd[k].detach().to(dev, torch.float32))
        self.lr = lr
        self.lam_rgb, self.lam_logp, self.logp_decay, self.lam_coh = rgb_loss_lambda, logp_loss_lambda, logp_loss_decay, loss_coh_lambda
        self.origin_step = origin_step
        self.frame_num = frame_num if frame_num is not None else latents.frame_num
        self.noise_std = sigma_noise_std
        self.seed = seed                   # int: jitter / sigma noise from the in-kernel Philox streams instead of torch's generator
        # the losses on the [N,3] maps and their gradients: two library kernels (tgtc_style_loss_sums / _grads) or, for
        # stand-in renderers and as the test oracle of those kernels, torch ops + autograd
        self.fused_losses = bool(fused_losses) and hasattr(renderer, "style_loss_sums") and self.lat.fused_ok()
        self.step_count = 0
        self.cnt = 0                       # the reference's `cnt` (train_tgtcs.py:347)
        self.prev = None                   # (x, y, x_origin): previous loss_coh batch's coarse / fine maps and its originals
        self._ws = [None, None]
        self.r.set_style_weights(*self.params)

    def _forward(self, rays_o, rays_d, lat, rand=None, slot=0):
        n = rays_o.shape[0]
        need = self.r.style_train_workspace_bytes(n)
        if self._ws[slot] is None or self._ws[slot].numel() < need:      # one persistent stash per pending batch
            self._ws[slot] = None
            self._ws[slot] = torch.empty(need, dtype=torch.uint8, device=self.r.device)
        if rand is None and self.seed is not None:
            rank = dist.get_rank(self.group) if dist.is_initialized() else 0
            key = (int(self.seed) * 0x9E3779B97F4A7C15 + (self.step_count << 24) + (rank << 12) + slot) & (2 ** 64 - 1)
            return self.r.style_train_forward(rays_o, rays_d, lat.detach(), workspace=self._ws[slot], seed=key, perturb=True,
                                              sigma_noise_std=self.noise_std)
        if rand is None:
            rand = torch.rand(n, 64, device=rays_o.device)           # perturb=True (train_tgtcs.py:362)
        nzc = nzf = None
        if self.noise_std > 0:
            nzc = torch.randn(n, 64, device=rays_o.device) * self.noise_std
            nzf = torch.randn(n, 128, device=rays_o.device) * self.noise_std
        return self.r.style_train_forward(rays_o, rays_d, lat.detach(), rand=rand, noise_coarse=nzc, noise_fine=nzf,
                                          workspace=self._ws[slot])

    def _coh_active(self):
        """the reference's `cnt` bookkeeping (train_tgtcs.py:397-404, :449-459): -> whether this iteration's coherence batch is
        compared with the previous one"""
        if self.cnt == self.frame_num:
            self.cnt = 1
            return False
        active = self.cnt != 0 and self.prev is not None
        self.cnt += 1
        return active

    def _step_fused(self, batch, coh_batch, gstep, world):
        """every per-sample and per-ray operation of the iteration is a library call; torch only concatenates the two batches"""
        dev = self.r.device
        sid, fid = batch["style_id"].long().to(dev).contiguous(), batch["frame_id"].long().to(dev).contiguous()
        gt = batch["rgb_gt"].contiguous()
        n = gt.shape[0]
        lam = self.lam_logp * (self.logp_decay ** int((gstep - self.origin_step) / 1000))             # train_tgtcs.py:426
        use_coh = gstep <= 122000                                    # train_tgtcs.py:486-493
        active = coh_batch is not None and self._coh_active()
        with_coh = use_coh and active
        coh = None
        if coh_batch is not None:
            sid2, fid2 = coh_batch["style_id"].long().to(dev).contiguous(), coh_batch["frame_id"].long().to(dev).contiguous()
            org2 = coh_batch["rgb_origin"].contiguous()
        if with_coh:
            # both batches of the iteration through ONE forward / backward (latents are per ray anyway): half the launches
            ids = (torch.cat([sid, sid2]), torch.cat([fid, fid2]))
            lat_all, logp_sum = self.lat.forward_fused(ids[0], ids[1], n)          # minus_logp covers the shuffled batch only
            ro = torch.cat([batch["rays_o"], coh_batch["rays_o"]])
            rd = torch.cat([batch["rays_d"], coh_batch["rays_d"]])
            rand = torch.cat([batch["rand"], coh_batch["rand"]]) if ("rand" in batch and "rand" in coh_batch) else None
            fw = self._forward(ro, rd, lat_all, rand)
            rc, rf = fw["rgb_coarse"], fw["rgb_fine"]
            rgb_c, rgb_f, c2, f2 = rc[:n], rf[:n], rc[n:], rf[n:]
            coh = (c2, f2, self.prev[0], self.prev[1], org2, self.prev[2])
        else:
            ids = (sid, fid)
            lat1, logp_sum = self.lat.forward_fused(sid, fid, n)
            fw = self._forward(batch["rays_o"], batch["rays_d"], lat1, batch.get("rand"))
            rgb_c, rgb_f = fw["rgb_coarse"], fw["rgb_fine"]
            if coh_batch is not None:                                # forward only: its maps are the next iteration's x / y
                lat2 = self.lat.forward_fused(sid2, fid2, 0)[0]
                fw2 = self._forward(coh_batch["rays_o"], coh_batch["rays_d"], lat2, coh_batch.get("rand"), slot=1)
                c2, f2 = fw2["rgb_coarse"], fw2["rgb_fine"]
                if active:
                    coh = (c2, f2, self.prev[0], self.prev[1], org2, self.prev[2])
        sums = self.r.style_loss_sums(rgb_c, rgb_f, gt, coh)        # [sq err coarse, sq err fine, coh ss coarse, coh ss fine]
        if world > 1:                                                # batch means and coherence norms are over all ranks' rows
            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=self.group)
        scale_rgb = self.lam_rgb / (3.0 * n * world)
        if with_coh:
            d_c, d_f = torch.empty_like(rc), torch.empty_like(rf)
            self.r.style_loss_grads(rgb_c, rgb_f, gt, scale_rgb, coh, sums[2:4], self.lam_coh, out=(d_c[:n], d_f[:n], d_c[n:], d_f[n:]))
        else:
            d_c, d_f, _, _ = self.r.style_loss_grads(rgb_c, rgb_f, gt, scale_rgb)
        bw = self.r.style_train_backward(fw["state"], d_c, d_f, grads=self.grads, accumulate=False)
        # latent table: the path through the style modules (d_latents) + the direct minus_logp term, one kernel.  Only the
        # shuffled batch's rows: latents_model_1.optimize(loss) backpropagates loss = loss_rgb + loss_logp
        # (train_tgtcs.py:481, :495; models.py:544-549), never the coherence term, which reaches the style modules only (:483-493)
        self.lat.backward_fused(sid, fid, n, bw["d_latents"][:n].contiguous(), lam / (n * world))
        if coh_batch is not None:
            self.prev = (c2, f2, org2)
        loss_rgb = self.lam_rgb * (sums[0] + sums[1]) / (3.0 * n * world)
        loss_coh = (torch.sqrt(sums[2] + 1e-8) + torch.sqrt(sums[3] + 1e-8)) if coh is not None else torch.zeros((), device=dev)
        loss_logp = lam * logp_sum[0] / (n * world)
        if world > 1:
            dist.all_reduce(self.grads, op=dist.ReduceOp.SUM, group=self.group)
            dist.all_reduce(self.lat.latents.grad, op=dist.ReduceOp.SUM, group=self.group)
            dist.all_reduce(loss_logp, op=dist.ReduceOp.SUM, group=self.group)
        loss = loss_rgb + loss_logp + ((self.lam_coh * loss_coh) if use_coh else 0.0)
        return loss, loss_rgb, loss_logp, loss_coh

    def _step_torch(self, batch, coh_batch, gstep, world):
        dev = self.r.device
        sid, fid = batch["style_id"].long().to(dev), batch["frame_id"].long().to(dev)
        lat1 = self.lat(sid, fid)                                    # [N,32], differentiable w.r.t. the table
        fw = self._forward(batch["rays_o"], batch["rays_d"], lat1, batch.get("rand"))
        rgb_c = fw["rgb_coarse"].requires_grad_(True)
        rgb_f = fw["rgb_fine"].requires_grad_(True)
        gt = batch["rgb_gt"]
        # every rank holds an equal shard of the iteration's two batches: batch means become (local mean) / world, and the
        # coherence norm is taken over all ranks' rows (one scalar all-reduce per term), so the summed gradients equal the
        # single-process ones
        loss_rgb = self.lam_rgb * (torch.mean((rgb_c - gt) ** 2) + torch.mean((rgb_f - gt) ** 2)) / world   # train_tgtcs.py:425, :480-481
        lam = self.lam_logp * (self.logp_decay ** int((gstep - self.origin_step) / 1000))             # train_tgtcs.py:426
        loss_logp = lam * self.lat.minus_logp(sid, fid, lat1) / world
        loss_coh = torch.zeros((), device=dev)       # value of the term (global)
        coh_local = None                             # this rank's differentiable share: its gradient is d(term)/d(local rows)
        fw2 = None
        if coh_batch is not None:
            sid2, fid2 = coh_batch["style_id"].long().to(dev), coh_batch["frame_id"].long().to(dev)
            lat2 = self.lat(sid2, fid2)
            fw2 = self._forward(coh_batch["rays_o"], coh_batch["rays_d"], lat2, coh_batch.get("rand"), slot=1)
            c2 = fw2["rgb_coarse"].requires_grad_(True)
            f2 = fw2["rgb_fine"].requires_grad_(True)
            org2 = coh_batch["rgb_origin"]
            # train_tgtcs.py:397-404, :449-459: compare with the previous batch unless a new pass over the frames starts
            if self._coh_active():
                x, y, x_org = self.prev
                # :401 compares with the previous originals; :456 runs after :403 replaced x_origin by THIS batch's originals
                terms = (cosine_similarity_rows(c2, x) - cosine_similarity_rows(org2, x_org),
                         cosine_similarity_rows(f2, y) - cosine_similarity_rows(org2, org2))
                coh_local = 0.0
                for v in terms:
                    ss = torch.sum(v ** 2)
                    ss_all = ss.detach().clone()
                    if world > 1:
                        dist.all_reduce(ss_all, op=dist.ReduceOp.SUM, group=self.group)
                    norm = torch.sqrt(ss_all + 1e-8)                # utils.L2_norm over all ranks' rows
                    loss_coh = loss_coh + norm
                    coh_local = coh_local + ss / (2.0 * norm)        # d/dv = v / norm = d L2_norm / dv
            self.prev = (c2.detach(), f2.detach(), org2)
        use_coh = gstep <= 122000                                    # train_tgtcs.py:486-493
        loss = loss_rgb + loss_logp + ((self.lam_coh * loss_coh) if use_coh else 0.0)
        objective = loss_rgb + loss_logp + ((self.lam_coh * coh_local) if (use_coh and coh_local is not None) else 0.0)
        # d loss / d (rgb maps) and the direct latent term (minus_logp) by torch on the tiny tensors
        self.lat.zero_grad()
        with_coh = use_coh and coh_local is not None
        leaves = [rgb_c, rgb_f] + ([c2, f2] if with_coh else [])
        gr = torch.autograd.grad(objective, leaves, retain_graph=True)
        bw = self.r.style_train_backward(fw["state"], gr[0], gr[1], grads=self.grads, accumulate=False)
        roots, seeds = [loss_logp, lat1], [None, bw["d_latents"]]    # direct term + the path through the style modules
        if with_coh:
            # the coherence batch's gradient reaches the two style modules only: latents_model_1.optimize(loss) backpropagates
            # loss = loss_rgb + loss_logp (train_tgtcs.py:481, :495), not loss_for_style
            self.r.style_train_backward(fw2["state"], gr[2], gr[3], grads=self.grads, accumulate=True)
        torch.autograd.backward(roots, seeds)                        # one engine run -> latents table
        if world > 1:
            dist.all_reduce(self.grads, op=dist.ReduceOp.SUM, group=self.group)
            dist.all_reduce(self.lat.latents.grad, op=dist.ReduceOp.SUM, group=self.group)
            for t in (loss_rgb, loss_logp):
                dist.all_reduce(t.detach_(), op=dist.ReduceOp.SUM, group=self.group)
            loss = loss_rgb + loss_logp + ((self.lam_coh * loss_coh) if use_coh else 0.0)
        return loss, loss_rgb, loss_logp, loss_coh

    def step(self, batch, coh_batch=None, global_step=None):
        """batch / coh_batch: dicts with rays_o, rays_d [N,3], rgb_gt [N,3], style_id, frame_id [N] (+ rgb_origin [N,3] in
        coh_batch), as LightDataLoader.get_batch / loss_coh_get_batch return them (train_tgtcs.py:356-370); an optional
        "rand" [N,64] replays the stratified-sampling uniforms (tests).
        Returns {"loss", "loss_rgb", "loss_logp", "loss_coh"} (device scalars)."""
        gstep = self.step_count if global_step is None else global_step
        dev = self.r.device
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        if self.fused_losses:
            loss, loss_rgb, loss_logp, loss_coh = self._step_fused(batch, coh_batch, gstep, world)
        else:
            loss, loss_rgb, loss_logp, loss_coh = self._step_torch(batch, coh_batch, gstep, world)
        self.step_count += 1
        self.r.adam_step(self.flat, self.grads, self.exp_avg, self.exp_avg_sq, self.step_count, lr=self.lr)   # style_optimizer (:54)
        self.lat.step()                                              # latents_model_1.optimize (:495)
        self.r.set_style_weights(*self.params)
        return {"loss": loss.detach(), "loss_rgb": loss_rgb.detach(), "loss_logp": loss_logp.detach(), "loss_coh": loss_coh.detach()}

    def state_dicts(self):
        return tuple({k: p.detach() for k, p in d.items()} for d in self.params)
