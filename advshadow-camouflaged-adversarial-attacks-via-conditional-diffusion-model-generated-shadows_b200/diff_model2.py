"""Drop-in replacement of the reference's `ddim2/diff_model2.py` (`from diff_model2 import *`,
ddim2/main2.py:6): the bigger UNet defaults (dm2:196-207), the linear schedule default (dm2:332),
the shadow helpers (dm2:457-654) -- plus `ddim_sample`, which the reference only ships in
diff_model.py:416-474 (additive; no caller breaks).
"""
from ._compat import *  # noqa: F401,F403
from ._compat import F, PILImage, models, torch  # noqa: F401
from ._diffusion import (GaussianDiffusionBase, cosine_beta_schedule, ddim_timestep_tables,  # noqa: F401
                         linear_beta_schedule)
from ._model import UNetModelBase
from . import shadow as _shadow
from .diff_model import norm_layer, timestep_embedding  # noqa: F401


class PretrainedResNet50:
    """Victim classifier wrapper (dm2:19-44).  Stays PyTorch, as the north star specifies."""

    def __init__(self, weight_path, device):
        self.model = models.resnet50(weights=None)
        self.model.load_state_dict(torch.load(weight_path, map_location=device))
        self.model = self.model.to(device)
        self.model.eval()

    def predict(self, images):
        with torch.no_grad():
            return self.model(images)


class UNetModel(UNetModelBase):
    def __init__(self, in_channels=3, model_channels=128, out_channels=3, num_res_blocks=3,
                 attention_resolutions=(4, 8, 16, 32), dropout=0.1, channel_mult=(1, 2, 4, 8),
                 conv_resample=True, num_heads=4):
        super().__init__(in_channels=in_channels, model_channels=model_channels, out_channels=out_channels,
                         num_res_blocks=num_res_blocks, attention_resolutions=attention_resolutions,
                         dropout=dropout, channel_mult=channel_mult, conv_resample=conv_resample,
                         num_heads=num_heads)


class GaussianDiffusion(GaussianDiffusionBase):
    _DEFAULT_SCHEDULE = 'linear'          # dm2:332

    def __init__(self, timesteps=1000, beta_schedule='linear'):
        super().__init__(timesteps, beta_schedule)

    # ---- shadow helpers ----
    def create_shadow_mask(self, image_size, shadow_center, shadow_radius, device):
        """[H,W] {0,1} disk mask (dm2:552-570)."""
        _, H, W = image_size
        c = shadow_center.detach().to(device).reshape(1, 2)
        r = torch.as_tensor(shadow_radius).detach().to(device).reshape(1)
        return _shadow.disk_mask(c, r, H, W)[0]

    def apply_adversarial_perturbation(self, classifier, image, target_label, device, epsilon=0.00001):
        """One FGSM step against the victim (dm2:572-613).  Victim forward/backward is PyTorch autograd."""
        victim = classifier.model
        x = image.detach().unsqueeze(0).to(device).clone().requires_grad_(True)
        with torch.enable_grad():
            loss = F.cross_entropy(victim(x), target_label)
            victim.zero_grad()
            loss.backward()
        return torch.clamp(x + epsilon * x.grad.data.sign(), 0, 1).detach()

    def apply_shadow(self, image, shadow_center, shadow_radius, feature_mask, classifier, target_label, device,
                     shadow_intensity=0.33, epsilon=0.01):
        """Shadow + in-mask adversarial perturbation (dm2:615-654); returns [1,C,H,W] like the reference
        (the perturbed image carries a leading batch axis that broadcasts into the result)."""
        image = image.to(device)
        feature_mask = feature_mask.to(device)
        Cc, H, W = image.shape
        sm = self.create_shadow_mask((Cc, H, W), shadow_center, shadow_radius, device)[None]
        img4, fm4 = image[None], feature_mask.reshape(1, -1, H, W)
        shadowed, _ = _shadow.composite(img4, sm, fm4, shadow_intensity, want_out=False)
        adv = self.apply_adversarial_perturbation(classifier, shadowed[0], target_label, device, epsilon)
        _, out = _shadow.composite(img4, sm, fm4, shadow_intensity, adv=adv, want_shadowed=False)
        return out

    def optimize_shadow_position(self, classifier, original_image, mask, target_label, device, lr=1e-1,
                                 iterations=10, verbose=False):
        """Adam on (centre, radius) (dm2:457-550).  As in the reference, the hard disk mask passes no
        gradient, so only the regulariser moves the parameters; the composite and the mask run on the
        GPU kernels, the victim stays PyTorch."""
        mask_center = _shadow.mask_center(mask)
        shadow_center = torch.nn.Parameter(mask_center.clone(), requires_grad=True)
        shadow_radius = torch.nn.Parameter(torch.tensor(20.0), requires_grad=True)
        original_image = original_image.to(device)
        mask = mask.to(device)
        optimizer = torch.optim.Adam([shadow_center, shadow_radius], lr=lr)
        victim = classifier.model.to(device)
        shadowed_image = None
        for iteration in range(iterations):
            optimizer.zero_grad()
            shadowed_image = self.apply_shadow(image=original_image, shadow_center=shadow_center,
                                               shadow_radius=shadow_radius, feature_mask=mask,
                                               classifier=classifier, target_label=target_label, device=device)
            img = shadowed_image.squeeze(0) if shadowed_image.dim() == 4 else shadowed_image
            adversarial_loss = -F.cross_entropy(victim(img.unsqueeze(0)), target_label)
            natural_loss = F.mse_loss(img, original_image)
            regularization = (shadow_center - mask_center).pow(2).sum() + shadow_radius.pow(2)
            loss = adversarial_loss + natural_loss + 0.1 * regularization
            loss.backward()
            if verbose:
                print(f"Iteration {iteration}: Loss={loss.item()}, adv_loss={adversarial_loss.item()}")
            if shadow_center.grad is not None and shadow_radius.grad is not None:
                optimizer.step()
            with torch.no_grad():
                shadow_center.clamp_(min=0, max=original_image.size(2))
                shadow_radius.clamp_(min=0, max=min(original_image.size(1), original_image.size(2)) / 2)
        return shadow_center.detach(), shadow_radius.detach(), shadowed_image

    def optimize_shadow_position_batched(self, classifier, original_images, masks, target_labels, device, lr=1e-1,
                                         iterations=10, shadow_intensity=0.33, epsilon=0.01):
        """`optimize_shadow_position` for B images at once -- the reference runs a Python loop with
        batch-1 victim calls per image (ddim2/main2.py:159-168).  Per image this is exactly the single-image algorithm
        (Adam is element-wise, the FGSM sign ignores the 1/B of the mean cross-entropy): masks and composites run on
        the batched GPU kernels, the victim forward/backward is one batched PyTorch call per iteration.
        original_images [B,C,H,W], masks [B,1|C,H,W], target_labels [B] -> (centers [B,2], radii [B], images [B,C,H,W])."""
        imgs = original_images.to(device).float()
        masks = masks.to(device).float()
        B, Cc, H, W = imgs.shape
        centers0 = torch.stack([_shadow.mask_center(masks[i]) for i in range(B)]).to(device)
        centers = torch.nn.Parameter(centers0.clone(), requires_grad=True)
        radii = torch.nn.Parameter(torch.full((B,), 20.0, device=device), requires_grad=True)
        optimizer = torch.optim.Adam([centers, radii], lr=lr)
        victim = classifier.model.to(device)
        out = None
        for _ in range(iterations):
            optimizer.zero_grad()
            sm = _shadow.disk_mask(centers, radii, H, W)
            shadowed, _ = _shadow.composite(imgs, sm, masks, shadow_intensity, want_out=False)
            x = shadowed.detach().clone().requires_grad_(True)          # FGSM step, dm2:572-613
            with torch.enable_grad():
                loss_adv = F.cross_entropy(victim(x), target_labels)
                victim.zero_grad()
                loss_adv.backward()
            adv = torch.clamp(x + epsilon * x.grad.data.sign(), 0, 1).detach()
            _, out = _shadow.composite(imgs, sm, masks, shadow_intensity, adv=adv, want_shadowed=False)
            # only the regulariser reaches (centre, radius): the hard mask passes no gradient (dm2:503-510)
            reg = (centers - centers0).pow(2).sum() + radii.pow(2).sum()
            (0.1 * reg).backward()
            optimizer.step()
            with torch.no_grad():
                centers.clamp_(min=0, max=W)
                radii.clamp_(min=0, max=min(H, W) / 2)
        return centers.detach(), radii.detach(), out

    def train_losses(self, model, x_start, t, device=None):   # dm2:656-679
        return super().train_losses(model, x_start, t)
