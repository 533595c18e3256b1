"""Seeded synthetic clouds used by the golden generator, the tests and the benchmark."""
import torch


def normalize(xyz: torch.Tensor) -> torch.Tensor:
    """Per-cloud centroid subtraction and division by the largest norm (the reference's provider.normalization)."""
    xyz = xyz - xyz.mean(dim=1, keepdim=True)
    scale = xyz.norm(dim=-1).max(dim=1)[0].view(-1, 1, 1)
    return xyz / scale


def uniform_cloud(B, N, seed):
    g = torch.Generator().manual_seed(seed)
    return normalize(torch.rand(B, N, 3, generator=g) * 2 - 1).float()


def surface_cloud(B, N, seed):
    """Points on the faces of a random cuboid, half of the box clipped away (cf. data_build/Cube.py)."""
    g = torch.Generator().manual_seed(seed)
    half = torch.rand(B, 1, 3, generator=g) * 0.8 + 0.2
    p = (torch.rand(B, N, 3, generator=g) * 2 - 1) * half
    axis = torch.randint(0, 3, (B, N), generator=g)
    side = torch.randint(0, 2, (B, N), generator=g).float() * 2 - 1
    onehot = torch.nn.functional.one_hot(axis, 3).bool()
    p = torch.where(onehot, side.unsqueeze(-1) * half.expand(B, N, 3), p)
    p[..., 0] = p[..., 0].abs()          # keep the visible half
    return normalize(p).float()


def duplicate_heavy(xyz, frac, seed):
    """Replace a fraction of the points by point 0 (the reference's provider.random_point_dropout)."""
    g = torch.Generator().manual_seed(seed)
    m = torch.rand(xyz.shape[:2], generator=g) < frac
    out = xyz.clone()
    out[m] = xyz[:, :1].expand(-1, xyz.shape[1], -1)[m]
    return out


def with_category(xyz, num_category=7):
    """xyz [B,N,3] -> [B,N,3+num_category] with the one-hot class of cloud b = b % num_category
    (the reference's provider.splice_torch)."""
    B, N, _ = xyz.shape
    cat = torch.zeros(B, N, num_category)
    for b in range(B):
        cat[b, :, b % num_category] = 1.0
    return torch.cat([xyz, cat], dim=-1)


def channels_first(x_bnc):
    """[B,N,C] -> the [B,C,N] contiguous tensor the models take; its .permute(0,2,1) is the strided layout."""
    return x_bnc.permute(0, 2, 1).contiguous()
