"""GAN trainer: the fast path of the reference's training loop (src/gan/train_gan.py:159-251).

Keeps the loop's semantics -- every batch one critic (D) step, every CRITIC_ITERS-th batch a generator
(G) step on the same batch, Adam(lr, betas) for D and for G+E_num, frozen eval-mode emotion
discriminator -- but runs each step body as one fused native call, keeps every parameter group in
flat buffers (one Adam launch, one gradient all-reduce per group), draws noise / alpha / dropout masks
with a counter-based device RNG, accumulates the losses on the device (no per-step .item()), and can
replay a whole cycle (CRITIC_ITERS D-steps + 1 G-step) as a single CUDA graph.

Data parallelism: one process per GPU, batch-sharded; the only exchange step is the
all-reduce of the flat gradient buffers over NCCL/NVLink (1.09 MB per D-step, 35.4 MB per G-step);
BatchNorm statistics stay local to the rank ("local BN", what torch DDP would do to the reference).
"""
import os

import torch
import torch.nn as nn

from . import _native
from . import dist as D_
from . import engine as E
from .optim import FlatParams, FusedAdam

_U64 = (1 << 64) - 1


def seed_everything(seed=42):
    """reference src/gan/utils.py:30-35"""
    import random
    import numpy as np
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def weights_init(m):
    """reference src/gan/utils.py:37-45: N(0, 0.02) weights and zero biases for every *Conv* / *Linear* module."""
    name = m.__class__.__name__
    if 'Conv' in name or 'Linear' in name:
        w = getattr(m, 'weight', None)
        if isinstance(w, torch.Tensor):
            nn.init.normal_(w.data, 0.0, 0.02)
        if getattr(m, 'bias', None) is not None:
            nn.init.constant_(m.bias.data, 0.0)


class GanTrainer:
    def __init__(self, cfg, ed_cfg, batch=None, precision="fp32", device=None, ed_state_dict=None,
                 process_group=None, seed_offset=0, modules=None, sync_bn=False, peer_allreduce=None):
        from src.gan.feature_encoder import FeatureEncoder
        from src.gan.models import Discriminator, Generator
        from src.emotion_discriminator.ed_model import EmotionDiscriminator

        if not torch.cuda.is_available():
            raise RuntimeError("GanTrainer needs a CUDA (sm_100a) device; there is no CPU fallback")
        self.cfg, self.ed_cfg = cfg, ed_cfg
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.B = int(batch if batch is not None else cfg.get('BATCH_SIZE', 32))
        self.precision = precision
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        self.critic_iters = int(cfg.get('CRITIC_ITERS', 5))
        emb_dim = cfg.get('ENCODER_OUT_DIM', 128)
        if modules is None:
            seed_everything(cfg.get("SEED", 42))
            # construction order and initialisation of train_gan.py:85-133
            E_num = FeatureEncoder(in_dim=cfg.get('NUMERIC_INPUT_DIM', 6), hidden_dims=cfg.get('ENCODER_HIDDEN', [256, 128]),
                                   out_dim=emb_dim)
            G = Generator(noise_dim=cfg['NOISE_DIM'], latent_dim=cfg['LATENT_DIM'],
                          mode=cfg.get('INTEGRATION_MODE', 'conditioning'), max_notes=cfg['MAX_NOTES'],
                          note_dim=cfg['NOTE_DIM'], numeric_embed_dim=emb_dim)
            D = Discriminator(max_notes=cfg['MAX_NOTES'], note_dim=cfg['NOTE_DIM'], numeric_embed_dim=emb_dim)
            ED = EmotionDiscriminator(ed_cfg)
            E_num.apply(weights_init); G.apply(weights_init); D.apply(weights_init)
            if ed_state_dict is not None:
                ED.load_state_dict(ed_state_dict, strict=False)
        else:
            E_num, G, D, ED = modules
        self.cond_dim = int(G.latent_dim) if G.mode == "conditioning" else 0
        self.E_num, self.G, self.D, self.ED = (m.to(self.device) for m in (E_num, G, D, ED))
        for p in self.ED.parameters():
            p.requires_grad = False
        self.ED.eval()
        self.G.train(); self.E_num.train(); self.D.train()

        # flat parameter groups in the reference's optimizer order (train_gan.py:136-145)
        self.flat_g = FlatParams(list(self.G.parameters()) + list(self.E_num.parameters()))
        self.flat_d = FlatParams(list(self.D.parameters()))
        betas = (cfg.get('BETA1', 0.5), cfg.get('BETA2', 0.9))
        self.opt_G = FusedAdam(self.flat_g, lr=float(cfg['LR_G']), betas=betas)
        self.opt_D = FusedAdam(self.flat_d, lr=float(cfg['LR_D']), betas=betas)
        self.opt_G.grad_scale = self.opt_D.grad_scale = 1.0 / self.world

        h = tuple(cfg.get('ENCODER_HIDDEN', [256, 128]))
        self.engine = E.GanEngine(self.B, precision=precision, max_notes=cfg['MAX_NOTES'], note_dim=cfg['NOTE_DIM'],
                                  noise_dim=cfg['NOISE_DIM'], latent_dim=cfg['LATENT_DIM'], gen_hidden=self.G.hidden,
                                  numeric_dim=cfg.get('NUMERIC_INPUT_DIM', 6), enc_hidden=h, embed_dim=emb_dim,
                                  n_classes=ed_cfg.get('n_classes', 4), enc_dropout=self.E_num.dropout,
                                  lambda_gp=cfg.get('LAMBDA_GP', 10.0), lambda_emotion=cfg.get('LAMBDA_EMOTION', 1.0),
                                  device=self.device, cond_dim=self.cond_dim)
        # 'conditioning' mode: the AE latents of the batch (encoder_feats.npy rows) live in one persistent buffer
        self.cond = torch.zeros((self.B, self.cond_dim), device=self.device) if self.cond_dim else None
        if self.cond is not None:
            self.engine.set_condition(self.cond)
        # data parallel: BatchNorm batch statistics are local to the rank by default (what torch DDP does to the reference);
        # sync_bn=True (config key SYNC_BN) makes them global, over NVLink peer memory (engine.sync_bn_connect)
        self.sync_bn = bool(sync_bn or cfg.get('SYNC_BN', False)) and self.world > 1
        if self.sync_bn:
            self.engine.sync_bn_connect(self.pg)
        self.rebind()

        dev, B = self.device, self.B
        self.noise = torch.empty((B, cfg['NOISE_DIM']), device=dev)
        self.alpha = torch.empty(B, device=dev)
        self.mask1 = torch.empty((B, h[0]), device=dev)
        self.mask2 = torch.empty((B, h[1]), device=dev)
        self.rng_seed = (int(cfg.get("SEED", 42)) * 0x9E3779B97F4A7C15 + seed_offset * 0xD1B54A32D192ED03) & _U64
        self.rng_counter = torch.zeros(1, dtype=torch.int64, device=dev)
        self.keep = 1.0 - self.E_num.dropout
        # device-side loss accumulators: [sum loss_d, sum gp, sum d_real, sum d_fake, sum g_adv, sum g_emo, n_d, n_g]
        self.loss_acc = torch.zeros(8, device=dev)
        self.m_d = torch.empty(4, device=dev)
        self.m_g = torch.empty(2, device=dev)
        self._graph = None
        # gradient exchange: NCCL all-reduce between two graphs per step (default), or kernels over NVLink peer memory
        # (peer_allreduce=True / MELOGAN_PEER_ALLREDUCE=1: no NCCL in the step, the whole data-parallel cycle is ONE graph)
        if peer_allreduce is None:
            peer_allreduce = os.environ.get("MELOGAN_PEER_ALLREDUCE") == "1"
        self._peer = None
        if peer_allreduce and self.world > 1:
            n = max(self.flat_g.grad.numel(), self.flat_d.grad.numel())
            self._peer = D_.PeerAllReduce(n, group=self.pg, device=self.device)

    # ---- binding ----
    def rebind(self):
        def views(module, keys, flat):
            named = dict(module.named_parameters()); named.update(dict(module.named_buffers()))
            P = {k: named[k].data for k in keys}
            G = {k: named[k].grad for k in keys if k in dict(module.named_parameters())}
            return P, G
        Pe, Ge = views(self.E_num, E.E_KEYS, self.flat_g)
        Pg, Gg = views(self.G, E.G_PARAM_KEYS + E.G_BUFFER_KEYS, self.flat_g)
        Pd, Gd = views(self.D, E.D_KEYS, self.flat_d)
        named = dict(self.ED.named_parameters()); named.update(dict(self.ED.named_buffers()))
        Ped = {k: named[k].data.contiguous() for k in E.ED_KEYS}
        self.engine.bind(E.MOD_E, Pe, Ge)
        self.engine.bind(E.MOD_G, Pg, Gg)
        self.engine.bind(E.MOD_D, Pd, Gd)
        self.engine.bind(E.MOD_ED, Ped, None)
        # parameters of this trainer change through FusedAdam.step only (which invalidates the packed copies it touches)
        self.engine.weight_cache(True)

    # ---- randomness ----
    def _draw(self, critic):
        """noise, (alpha), mask1, mask2 for one step from the counter-based device RNG."""
        L, st = _native.lib(), self.engine._stream()
        ctr = self.rng_counter.data_ptr()
        jobs = [(self.noise, 0, 0.0, 1), (self.mask1, 2, self.keep, 2), (self.mask2, 2, self.keep, 3)]
        if critic:
            jobs.append((self.alpha, 1, 0.0, 4))
        with torch.cuda.device(self.device):
            for t, kind, p, salt in jobs:
                _native.check(L.mg_rng_fill_counter(t.data_ptr(), t.numel(), kind, p, (self.rng_seed + salt) & _U64, ctr,
                                                    1 << 22, st))
            _native.check(L.mg_counter_add(ctr, 1, st))

    def _allreduce(self, flat):
        if self._peer is not None:
            self._peer.allreduce_sum_(flat.grad)       # three kernels over NVLink peer memory (graph-capturable)
        elif self.world > 1:
            D_.allreduce_sum_(flat.grad, self.pg)      # the 1/world factor is folded into Adam's grad_scale

    # ---- the two step bodies ----
    def _set_cond(self, cond):
        if self.cond_dim:
            if cond is None:
                raise ValueError("INTEGRATION_MODE 'conditioning' needs the AE latents of the batch (encoder_feats.npy)")
            self.cond.copy_(cond)

    def critic_step(self, real, numeric, noise=None, alpha=None, mask1=None, mask2=None, cond=None):
        """train_gan.py:183-205.  Returns the device tensor [loss_d, gp, mean D(real), mean D(fake)]."""
        self._set_cond(cond)
        if noise is None:
            self._draw(critic=True)
            noise, alpha, mask1, mask2 = self.noise, self.alpha, self.mask1, self.mask2
        self.opt_D.zero_grad()
        self.engine.critic_step(real, numeric, noise, alpha, mask1, mask2, metrics=self.m_d)
        for bn in (self.G.decoder.deconv[1], self.G.decoder.deconv[4]):
            bn.num_batches_tracked += 1
        self._allreduce(self.flat_d)
        self.opt_D.step()
        self.loss_acc[0:4] += self.m_d
        self.loss_acc[6] += 1
        return self.m_d

    def generator_step(self, numeric, labels, noise=None, mask1=None, mask2=None, cond=None):
        """train_gan.py:212-251.  Returns the device tensor [loss_g_adv, loss_g_emo]."""
        self._set_cond(cond)
        if noise is None:
            self._draw(critic=False)
            noise, mask1, mask2 = self.noise, self.mask1, self.mask2
        self.opt_G.zero_grad()
        self.engine.generator_step(numeric, noise, labels, mask1, mask2, metrics=self.m_g)
        for bn in (self.G.decoder.deconv[1], self.G.decoder.deconv[4]):
            bn.num_batches_tracked += 1
        self._allreduce(self.flat_g)
        self.opt_G.step()
        self.loss_acc[4:6] += self.m_g
        self.loss_acc[7] += 1
        return self.m_g

    def train_cycle(self, reals, numerics, labels, conds=None):
        """CRITIC_ITERS critic steps on reals[i], numerics[i], then one generator step on the last batch.
        reals (K, B, T, 4), numerics (K, B, F), labels (B,), conds (K, B, latent; 'conditioning' mode) are CUDA tensors."""
        K = self.critic_iters
        for i in range(K):
            self.critic_step(reals[i], numerics[i], cond=None if conds is None else conds[i])
        self.generator_step(numerics[K - 1], labels, cond=None if conds is None else conds[K - 1])

    # ---- whole-cycle CUDA graph ----
    def capture_cycle(self):
        """Captures train_cycle over static input buffers; returns them.  Run at least one eager cycle first
        (first calls allocate scratch, which capture forbids).
        In 'conditioning' mode a fourth static buffer, self.s_conds (K, B, latent), holds the AE latents of the batches.
        Single GPU: the whole cycle is ONE graph.  Data parallel: NCCL all-reduces stay outside the graphs (capturing
        them deadlocked on this stack), so every step is two graphs -- [zero grads, draw, step body] and [Adam, loss
        accumulation] -- with the eager all-reduce of the flat gradient buffer between them on the same stream.
        (Round 2, two B200s: NCCL captured inside ONE cycle graph in thread-local capture mode replays correctly but is
        no faster -- 53.70 vs 53.47 ms per cycle -- and the process group then hangs at teardown; not kept.)"""
        K, B, dev = self.critic_iters, self.B, self.device
        self.s_reals = torch.zeros((K, B, self.cfg['MAX_NOTES'], self.cfg['NOTE_DIM']), device=dev)
        self.s_numerics = torch.zeros((K, B, self.cfg.get('NUMERIC_INPUT_DIM', 6)), device=dev)
        self.s_labels = torch.zeros(B, dtype=torch.int64, device=dev)
        # 'conditioning' mode: the AE latents of the K batches are a fourth static input (self.s_conds); every captured
        # step copies its slice into the persistent buffer the native context reads
        self.s_conds = torch.zeros((K, B, self.cond_dim), device=dev) if self.cond_dim else None
        torch.cuda.synchronize(dev)
        if self.world == 1 or self._peer is not None:      # no NCCL in the step: the whole cycle is ONE graph
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.train_cycle(self.s_reals, self.s_numerics, self.s_labels, conds=self.s_conds)
            self._graph = g
            self._one_graph = True
            return self.s_reals, self.s_numerics, self.s_labels
        pool = torch.cuda.graph_pool_handle()
        self._g_pre, self._g_post = [], []
        bns = (self.G.decoder.deconv[1], self.G.decoder.deconv[4])
        for i in range(K + 1):
            pre, post = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(pre, pool=pool):
                if self.s_conds is not None:
                    self._set_cond(self.s_conds[min(i, K - 1)])
                if i < K:
                    self._draw(critic=True)
                    self.opt_D.zero_grad()
                    self.engine.critic_step(self.s_reals[i], self.s_numerics[i], self.noise, self.alpha, self.mask1,
                                            self.mask2, metrics=self.m_d)
                else:
                    self._draw(critic=False)
                    self.opt_G.zero_grad()
                    self.engine.generator_step(self.s_numerics[K - 1], self.noise, self.s_labels, self.mask1, self.mask2,
                                               metrics=self.m_g)
                for bn in bns:
                    bn.num_batches_tracked += 1
            with torch.cuda.graph(post, pool=pool):
                if i < K:
                    self.opt_D.step()
                    self.loss_acc[0:4] += self.m_d
                    self.loss_acc[6] += 1
                else:
                    self.opt_G.step()
                    self.loss_acc[4:6] += self.m_g
                    self.loss_acc[7] += 1
            self._g_pre.append(pre); self._g_post.append(post)
        self._graph = True
        return self.s_reals, self.s_numerics, self.s_labels

    def replay_cycle(self):
        if self.world == 1 or getattr(self, "_one_graph", False):
            self._graph.replay()
            return
        K = self.critic_iters
        for i in range(K + 1):
            self._g_pre[i].replay()
            self._allreduce(self.flat_d if i < K else self.flat_g)
            self._g_post[i].replay()

    # ---- host -> device input pipeline (double buffered) ----
    def prefetch(self, h_reals, h_numerics, h_labels, h_conds=None):
        """Starts the host->device copy of the NEXT cycle's inputs (pinned host tensors shaped like the static buffers of
        capture_cycle) on a dedicated copy stream into a second set of device buffers, so that it runs under the cycle
        the compute stream is working on.  replay_cycle_prefetched() then moves them into the static buffers with one
        device-to-device copy (337 MB at HBM speed instead of PCIe speed on the critical path)."""
        if self._graph is None:
            raise RuntimeError("capture_cycle() first")
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(self.device)
            self._stg = [torch.empty_like(t) for t in (self.s_reals, self.s_numerics, self.s_labels)]
            self._stg_conds = torch.empty_like(self.s_conds) if self.s_conds is not None else None
            self._staged, self._stg_free = torch.cuda.Event(), torch.cuda.Event()
            self._stg_free.record(torch.cuda.current_stream(self.device))
        cs = self._copy_stream
        cs.wait_event(self._stg_free)                    # the previous contents have left the staging buffers
        with torch.cuda.stream(cs):
            for dst, src in zip(self._stg, (h_reals, h_numerics, h_labels)):
                dst.copy_(src, non_blocking=True)
            if self._stg_conds is not None:
                if h_conds is None:
                    raise ValueError("INTEGRATION_MODE 'conditioning' needs the AE latents of the batches")
                self._stg_conds.copy_(h_conds, non_blocking=True)
            self._staged.record(cs)

    def replay_cycle_prefetched(self):
        """Replays the captured cycle on the inputs of the last prefetch()."""
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self._staged)
        self.s_reals.copy_(self._stg[0]); self.s_numerics.copy_(self._stg[1]); self.s_labels.copy_(self._stg[2])
        if self._stg_conds is not None:
            self.s_conds.copy_(self._stg_conds)
        self._stg_free.record(cur)
        self.replay_cycle()

    def sync_bn_running_stats_(self):
        """Data parallel: BatchNorm batch statistics are local to the rank (what torch DDP does to the reference), so
        the running estimates drift apart.  Before a checkpoint / evaluation they are replaced by their mean over ranks
        (the running mean of rank-local batch means IS the global-batch mean; for the variance the mean of the
        rank-local unbiased variances, torch SyncBatchNorm's estimate minus the between-rank term)."""
        if self.world > 1 and not self.sync_bn:          # (with SyncBatchNorm every rank already holds the same estimates)
            for bn in (self.G.decoder.deconv[1], self.G.decoder.deconv[4]):
                for t in (bn.running_mean, bn.running_var):
                    torch.distributed.all_reduce(t, group=self.pg)
                    t.div_(self.world)

    def epoch_means(self):
        """(D_loss, G_adv, G_emo) accumulated since the last call, one host sync (train_gan.py:254-264 log line).
        Data parallel: every rank's loss is the mean over its shard, so the sums are all-reduced first -- the log line
        is the mean over the global batch, not rank 0's shard."""
        if self.world > 1:
            D_.allreduce_sum_(self.loss_acc, self.pg)
        a = self.loss_acc.cpu()
        self.loss_acc.zero_()
        nd, ng = max(a[6].item(), 1.0), max(a[7].item(), 1.0)
        return a[0].item() / nd, a[4].item() / ng, a[5].item() / ng

    def state_dict(self):
        """Checkpoint layout of train_gan.py:269-276 (collective in data-parallel mode: every rank calls it)."""
        self.sync_bn_running_stats_()
        return {'G': self.G.state_dict(), 'D': self.D.state_dict(), 'E_num': self.E_num.state_dict(),
                'opt_G': self.opt_G.state_dict(), 'opt_D': self.opt_D.state_dict(),
                'rng_counter': self.rng_counter.clone()}

    def load_state_dict(self, ck):
        """Resume from a checkpoint of train_gan.py:269-276 (this trainer's or the reference's: same keys and optimizer
        state layout).  Parameters are copied INTO the flat groups, so the native bindings stay valid; `rng_counter`
        (ours only) restores the device noise stream, a reference checkpoint just starts a fresh one."""
        self.G.load_state_dict(ck['G'])
        self.D.load_state_dict(ck['D'])
        self.E_num.load_state_dict(ck['E_num'])
        if 'opt_G' in ck:
            self.opt_G.load_state_dict(ck['opt_G'])
        if 'opt_D' in ck:
            self.opt_D.load_state_dict(ck['opt_D'])
        if 'rng_counter' in ck:
            self.rng_counter.copy_(ck['rng_counter'].to(self.rng_counter.device))
        self.rebind()                      # also invalidates the packed-weight cache (mg_gan_bind)
