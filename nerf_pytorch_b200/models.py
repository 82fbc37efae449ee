"""Parameter container for the network family the fused path renders (``FlexibleNeRFModel``, nerf/models.py:185-256).

The render path never evaluates a network in PyTorch: ``train_utils`` reads the parameters of whatever module it is
handed (the reference's own class works as-is, it is duck-typed on its attribute names) and gives them to the CUDA
kernels.  This class exists for callers that do not import the reference: it builds the same *state_dict schema*
(names, shapes, ``nn.Linear`` default init in the same order, so a shared seed gives the reference's weights and
checkpoints load either way, train_nerf.py:373-383) from a table, and nothing else.

``forward`` (an encoded ``(N, dim_xyz + dim_dir)`` batch -> ``(N, 4)``) is provided for completeness and runs the
same table on plain ``torch.nn.functional.linear``; the skip input joins a layer exactly when that layer was built
wide, which is the behaviour the reference intends (its own forward raises on any live skip)."""
from __future__ import annotations

import torch
from torch import nn


def linear_table(num_layers, hidden, skip_every, dim_xyz, dim_dir, use_viewdirs):
    """[(attribute path, in_features, out_features)] in the reference's construction order (= its RNG order)."""
    rows = [("layer1", dim_xyz, hidden)]
    for i in range(num_layers - 1):
        wide = i > 0 and i % skip_every == 0 and i != num_layers - 1
        rows.append((f"layers_xyz.{i}", hidden + (dim_xyz if wide else 0), hidden))
    if use_viewdirs:
        rows += [("layers_dir.0", hidden + dim_dir, hidden // 2), ("fc_alpha", hidden, 1), ("fc_rgb", hidden // 2, 3),
                 ("fc_feat", hidden, hidden)]
    else:
        rows.append(("fc_out", hidden, 4))
    return rows


class FlexibleNeRFModel(nn.Module):
    def __init__(self, num_layers=4, hidden_size=128, skip_connect_every=4, num_encoding_fn_xyz=6, num_encoding_fn_dir=4,
                 include_input_xyz=True, include_input_dir=True, use_viewdirs=True):
        super().__init__()
        self.dim_xyz = 6 * num_encoding_fn_xyz + (3 if include_input_xyz else 0)
        self.dim_dir = (6 * num_encoding_fn_dir + (3 if include_input_dir else 0)) if use_viewdirs else 0
        self.skip_connect_every = skip_connect_every
        self.use_viewdirs = use_viewdirs
        for path, fin, fout in linear_table(num_layers, hidden_size, skip_connect_every, self.dim_xyz, self.dim_dir,
                                            use_viewdirs):
            owner, _, _ = path.partition(".")
            lin = nn.Linear(fin, fout)
            if owner in ("layers_xyz", "layers_dir"):
                if not hasattr(self, owner):
                    setattr(self, owner, nn.ModuleList())
                getattr(self, owner).append(lin)
            else:
                setattr(self, owner, lin)
        if not hasattr(self, "layers_xyz"):
            self.layers_xyz = nn.ModuleList()

    def forward(self, x):
        F = torch.nn.functional
        xyz, view = x[..., :self.dim_xyz], x[..., self.dim_xyz:]
        h = self.layer1(xyz)
        for lin in self.layers_xyz:
            h = F.relu(lin(torch.cat((h, xyz), -1) if lin.in_features > h.shape[-1] else h))
        if not self.use_viewdirs:
            return self.fc_out(h)
        sigma = self.fc_alpha(h)
        d = torch.cat((F.relu(self.fc_feat(h)), view), -1)
        for lin in self.layers_dir:
            d = F.relu(lin(d))
        return torch.cat((self.fc_rgb(d), sigma), -1)
