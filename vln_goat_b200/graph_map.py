"""Device-resident node-embedding bank of the navigation graph.

The reference keeps, per episode, a Python dict ``{viewpoint: [sum of embeddings, count]}`` of torch tensors and walks it
with string keys at every step (M/models/graph_utils.py:110-121 ``GraphMap.update_node_embed`` / ``get_node_embed``,
called from M/r2r/agent.py:560-571 and :163-171: the visited node gets its fused panorama embedding with ``rewrite=True``,
every candidate view adds its embedding to the neighbour it leads to, and a node's feature is ``sum / count``).  Here the
whole batch of episodes shares two tensors -- ``sum [episodes, max_nodes, H]`` and ``count [episodes, max_nodes]`` -- and
the per-step updates / reads are index ops over (episode, node) pairs that the agent already has as integers: no Python
loop over nodes, no host round trip, differentiable like the reference's in-place tensor sums."""
import torch


class NodeEmbedBank(object):
    def __init__(self, episodes, max_nodes, hidden, device="cuda", dtype=torch.float32):
        self.sum = torch.zeros(episodes, max_nodes, hidden, device=device, dtype=dtype)
        self.count = torch.zeros(episodes, max_nodes, device=device, dtype=dtype)

    def update(self, episode_idx, node_idx, embeds, rewrite=False):
        """embeds [n, H] for the (episode_idx[n], node_idx[n]) pairs.  rewrite=True: sum = embed, count = 1 (the visited node,
        graph_utils.py:111-112); else sum += embed, count += 1 (candidate views; duplicate pairs accumulate, :114-119).
        Out-of-place on purpose: earlier reads stay valid for autograd, as with the reference's fresh list entries."""
        e = episode_idx.to(torch.int64)
        n = node_idx.to(torch.int64)
        if rewrite:
            keep = torch.ones_like(self.count)
            keep[e, n] = 0.0
            self.sum = self.sum * keep.unsqueeze(-1)
            self.count = self.count * keep
        self.sum = self.sum.index_put((e, n), embeds.to(self.sum.dtype), accumulate=True)
        self.count = self.count.index_put((e, n), torch.ones_like(e, dtype=self.count.dtype), accumulate=True)

    def get(self, episode_idx, node_idx):
        """sum / count of the named nodes (graph_utils.py:121) -> [n, H]"""
        e = episode_idx.to(torch.int64)
        n = node_idx.to(torch.int64)
        return self.sum[e, n] / self.count[e, n].clamp(min=1.0).unsqueeze(-1)

    def get_padded(self, node_index_table):
        """node_index_table int [episodes, G] (-1 = empty slot) -> the [episodes, G, H] ``gmap_img_embeds`` block of a
        navigation batch (M/r2r/agent.py:163-171), zeros in empty slots."""
        t = node_index_table.to(torch.int64)
        valid = t >= 0
        e = torch.arange(t.shape[0], device=t.device).unsqueeze(1).expand_as(t)
        n = t.clamp(min=0)
        out = self.sum[e, n] / self.count[e, n].clamp(min=1.0).unsqueeze(-1)
        return out * valid.unsqueeze(-1).to(out.dtype)
