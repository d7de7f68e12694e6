"""Renderer: flattens a ray bundle, renders it and restores the image shape
(reference: models/renderer.py:6-65).

The reference loops over ``ray_chunk``-sized chunks in Python (313 iterations for an
800x800 frame).  Here the whole bundle goes to the CUDA library in one call; the chunk
size is still honoured where it is observable (the per-chunk inside-box predicate of the
sampler and the per-chunk random draws, see ``field.render_rays``)."""
import torch
import torch.nn as nn


class Renderer(nn.Module):
    def __init__(self, tensorf, batch_size, test_batch_size, ray_chunk, distance_scale=1, lindisp=False,
                 perturb=True, tensorf_sample=True, ndc=False):
        super().__init__()
        self.tensorf = tensorf
        self.batch_size, self.test_batch_size = batch_size, test_batch_size
        self.lindisp, self.perturb = lindisp, perturb
        self.distance_scale = distance_scale
        self.tensorf_sample = tensorf_sample
        self.ndc = ndc
        self.ray_chunk = ray_chunk

    def forward(self, t, rays, white_background=False, transfer_vel=False):
        if self.ndc:
            raise NotImplementedError("nvfi_b200: ndc rays are unsupported (renderer.ndc is False in every config)")
        ray_o = rays.ray_origins.reshape(-1, 3)
        ray_d = rays.ray_directions.reshape(-1, 3)
        field = getattr(self.tensorf, "nvfi", None)
        if field is not None and hasattr(field, "render_rays"):
            rgb, depth, acc, w, extra = field.render_rays(t, ray_o, ray_d, white_bg=white_background,
                                                          transfer_vel=transfer_vel, ray_chunk=self.ray_chunk)
        else:   # any object exposing the reference's per-chunk API
            outs = [[] for _ in range(5)]
            n = ray_o.shape[0]
            for c in range(n // self.ray_chunk + int(n % self.ray_chunk > 0)):
                sl = slice(c * self.ray_chunk, (c + 1) * self.ray_chunk)
                fn = self.tensorf.render_ray_transfer if transfer_vel else self.tensorf.render_ray
                for k, v in enumerate(fn(t, ray_o[sl], ray_d[sl], white_background, self.ndc)):
                    outs[k].append(v)
            rgb, depth, acc, w, extra = (torch.cat(x, 0) for x in outs)
        shape = tuple(rays.restore_shape)
        # the reference reshapes the 5th output to (..., 3) and therefore fails for
        # mask_dim != 3; reshape to its own width instead
        return (rgb.reshape(*shape, 3), depth.reshape(*shape), acc.reshape(*shape),
                w.reshape(*shape, w.shape[-1]), extra.reshape(*shape, extra.shape[-1]))

    def render(self, t, rays, white_background=False, mode="train", transfer_vel=False):
        if mode == "train":
            self.tensorf.train()
            return self.forward(t, rays, white_background)
        self.tensorf.eval()
        with torch.no_grad():
            return self.forward(t, rays, white_background, transfer_vel=transfer_vel)
