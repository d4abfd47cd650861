"""Drop-in ``RelGATModel`` (reference relgat_projector/core/model/model.py:13-292): frozen node
embeddings -> L RelGAT layers (+ELU between) -> optional ProjectionHead -> KG scorer, with the
GAT stack and the gather-score fused over the sm_100a kernels."""
from __future__ import annotations

import json
import os
from typing import Any, Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import functional as RF
from . import ops
from .blocks import build_blocks
from .graph import get_graph_index
from .layer import RelGATLayer
from .projection import ProjectionHead
from .scorer import DistMultScorer, TransEScorer


class RelGATModel(nn.Module):
    def __init__(
        self,
        node_emb: torch.Tensor,  # [N, D_in] frozen
        edge_index: torch.Tensor,  # [2, E]
        edge_type: torch.Tensor,  # [E]
        num_rel: int,
        scorer_type: str = "distmult",
        gat_out_dim: int = 200,
        gat_heads: int = 4,
        dropout: float = 0.2,
        relation_attn_dropout: float = 0.0,
        gat_num_layers: int = 1,
        project_to_input_size: bool = False,
        projection_layers: int = 1,
        projection_dropout: float = 0.0,
        projection_hidden_dim: int = 0,
        precision: str = "fp32",
    ):
        super().__init__()
        self.register_buffer("node_emb_fixed", node_emb)  # buffer, saved in state_dict (model.py:32)
        self.edge_index = edge_index  # plain attributes like the reference (model.py:34-35)
        self.edge_type = edge_type
        self.num_rel = num_rel
        self.gat_num_layers = gat_num_layers
        self.projection_layers = projection_layers
        self.project_to_input_size = project_to_input_size
        self.precision = precision
        # extension (not a constructor argument of the reference): run the batch-row path (forward / batch_rows) on the
        # batch's receptive-field blocks instead of the whole graph — same rows, same gradients, less work (blocks.py);
        # single_gat_step / evaluation always run the full graph
        self.receptive_field = os.environ.get("RELGAT_RECEPTIVE_FIELD", "0") != "0"
        # how: "masked" (default; fp32 mode) = the full graph's index with per-step work tables over the needed
        # destinations and compact source rows; "blocks" = per-step bipartite sub-indexes (blocks.py; any precision)
        self.receptive_field_mode = os.environ.get("RELGAT_RF_MODE", "masked")
        self.last_block_edges = None
        if project_to_input_size and self.projection_layers < 1:
            raise ValueError("projection_layers must be >= 1 when project_to_input_size=True")
        self._config = dict(
            input_dim=int(node_emb.size(1)), num_rel=int(num_rel), scorer_type=str(scorer_type),
            gat_out_dim=int(gat_out_dim), gat_heads=int(gat_heads), dropout=float(dropout),
            relation_attn_dropout=float(relation_attn_dropout), gat_num_layers=int(gat_num_layers),
            project_to_input_size=bool(project_to_input_size), projection_layers=int(projection_layers),
            projection_dropout=float(projection_dropout), projection_hidden_dim=int(projection_hidden_dim),
        )

        def make_layer(in_dim):
            return RelGATLayer(in_dim=in_dim, out_dim=gat_out_dim, num_rel=num_rel, heads=gat_heads,
                               dropout=dropout, relation_attn_dropout=relation_attn_dropout, use_bias=True,
                               precision=precision)

        if gat_num_layers == 1:  # attribute names follow the reference (model.py:44-73)
            self.gat_layer = make_layer(node_emb.size(1))
            self.act = None
        else:
            self.gat_layers = nn.ModuleList()
            in_dim = node_emb.size(1)
            for _ in range(max(1, gat_num_layers)):
                self.gat_layers.append(make_layer(in_dim))
                in_dim = gat_out_dim * gat_heads
            self.act = nn.ELU()

        scorer_dim = gat_out_dim * gat_heads
        if self.project_to_input_size:
            self.projection = ProjectionHead(in_dim=scorer_dim, out_dim=node_emb.size(1),
                                             hidden_dim=projection_hidden_dim, num_layers=self.projection_layers,
                                             dropout=projection_dropout, precision=precision)
            scorer_dim = node_emb.size(1)
        else:
            self.projection = None

        self.scorer_type = scorer_type
        if scorer_type.lower() == "distmult":
            self.scorer = DistMultScorer(num_rel, rel_dim=scorer_dim)
        elif scorer_type.lower() == "transe":
            self.scorer = TransEScorer(num_rel, rel_dim=scorer_dim, normalize=True)
        else:
            raise ValueError(f"Unknown scorer_type: {scorer_type}")
        self._x0_planes = None  # cached bf16 split of the frozen input embeddings

    # -- helpers ------------------------------------------------------------------------------
    def _layers(self) -> List[RelGATLayer]:
        return [self.gat_layer] if self.gat_num_layers == 1 else list(self.gat_layers)

    def _input_planes(self):
        x = self.node_emb_fixed
        key = (x.data_ptr(), x._version, tuple(x.shape), self.precision)
        if self._x0_planes is None or self._x0_planes[0] != key:
            self._x0_planes = (key, ops.split_bf16(x, with_lo=(self.precision == "fp32")))
        return self._x0_planes[1]

    def _graph(self):
        return get_graph_index(self.edge_index, self.edge_type, self.node_emb_fixed.size(0), self.num_rel)

    def _stack_output(self, gather_ids: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Output of the GAT stack before the projection head: all N rows, or only ``x[gather_ids]``.

        One autograd node for the whole stack: ELU between layers (reference model.py:286-287) and both dropouts of
        every layer (layer.py:296-297, 321-322) are applied inside the edge kernels, so training with the
        reference's default rates runs the same fused path as dropout 0."""
        layers = self._layers()
        graph = self._graph()
        dev = self.node_emb_fixed.device
        masked = self.receptive_field and self.receptive_field_mode == "masked" and self.precision == "fp32"
        if gather_ids is not None and self.receptive_field and not masked:
            # one bipartite block per layer, holding the batch's L-hop in-neighbourhood only
            blk = build_blocks(graph, gather_ids, len(layers))
            self.last_block_edges = blk.n_edges
            drop = [lyr.draw_dropout(g.N, g.E, dev) for lyr, g in zip(layers, blk.graphs)]
            drop = [d if d is not None else RF.LayerDropout() for d in drop] if any(d is not None for d in drop) else None
            planes = tuple(None if p is None else ops.gather_plane_rows(p, blk.input_rows) for p in self._input_planes())
            return RF.relgat_stack(None, blk.graphs, layers[0].heads, layers[0].out_dim,
                                   [lyr.kernel_params() for lyr in layers], precision=self.precision,
                                   x0_planes=planes, drop=drop, gather_ids=blk.out_pos)
        drop = [lyr.draw_dropout(graph.N, graph.E, dev) for lyr in layers]
        drop = [d if d is not None else RF.LayerDropout() for d in drop] if any(d is not None for d in drop) else None
        rows = RF.relgat_stack(self.node_emb_fixed, graph, layers[0].heads, layers[0].out_dim,
                               [lyr.kernel_params() for lyr in layers], precision=self.precision,
                               x0_planes=self._input_planes(), drop=drop, gather_ids=gather_ids,
                               prune=masked and gather_ids is not None)
        if masked and gather_ids is not None:
            self.last_block_edges = RF.LAST_PRUNED_EDGES
        return rows

    # -- reference API ------------------------------------------------------------------------
    def single_gat_step(self) -> torch.Tensor:
        """Full-graph node representations, optionally projected (reference model.py:274-292)."""
        x = self._stack_output()
        if self.project_to_input_size:
            x = self.projection(x)
        return x

    def _projection_dropout_active(self) -> bool:
        return (self.training and self.project_to_input_size
                and isinstance(self.projection.dropout, torch.nn.Dropout) and self.projection.dropout.p > 0.0)

    def batch_rows(self, ids: torch.Tensor) -> torch.Tensor:
        """``single_gat_step()[ids]`` without building all N projected rows or a dense [N, D] gradient: the stack
        returns only the named rows (its backward scatters their gradients into a persistent zero table) and the
        projection head — a row-wise map — runs on those rows.  Same values as the reference's order of operations
        (model.py:135-137); with an active PROJECTION dropout the mask would be drawn per gathered row instead of per
        node, so that case projects all rows first like the reference."""
        if self._projection_dropout_active():
            return self.single_gat_step()[ids]
        rows = self._stack_output(gather_ids=ids)
        return self.projection(rows) if self.project_to_input_size else rows

    def forward(self, src_ids: torch.Tensor, rel_ids: torch.Tensor, dst_ids: torch.Tensor,
                transform_to_input_if_possible: bool = True, transform_rows: Optional[int] = None
                ) -> Tuple[torch.Tensor, Optional[torch.Tensor], Optional[torch.Tensor]]:
        """Scores for a batch of triples (reference model.py:99-142): returns
        (scores [B], transformed [B, D_sc] or None, dst_vec [B, D_sc]).  ``transform_rows`` (extension): compute
        ``transformed`` for the first that many triples only (the trainer uses the positives)."""
        want_tr = self.project_to_input_size and transform_to_input_if_possible
        b = int(src_ids.numel())
        rows = self.batch_rows(torch.cat([src_ids, dst_ids]))
        src_vec, dst_vec = rows[:b], rows[b:]
        n_tr = (b if transform_rows is None else min(int(transform_rows), b)) if want_tr else 0
        scores, transformed = self.scorer.score_and_transform(src_vec, rel_ids, dst_vec, n_tr)
        return scores, (transformed if want_tr else None), dst_vec

    @torch.no_grad()
    def get_node_repr(self) -> torch.Tensor:
        return self.single_gat_step()

    @torch.no_grad()
    def transform(self, src_ids: torch.Tensor, rel_ids: torch.Tensor) -> torch.Tensor:
        x = self.single_gat_step()
        return self.transform_from_vectors(src_vectors=x[src_ids], rel_ids=rel_ids)

    @torch.no_grad()
    def transform_from_vectors(self, src_vectors: torch.Tensor, rel_ids: torch.Tensor) -> torch.Tensor:
        if rel_ids.dim() == 0:
            rel_ids = rel_ids.view(1)
        if rel_ids.numel() == 1 and src_vectors.size(0) > 1:
            rel_ids = rel_ids.expand(src_vectors.size(0))
        return self.scorer.transform(src_vectors, rel_ids)

    def get_config(self) -> dict:
        """Constructor arguments needed to rebuild the model (the reference reads an attribute it
        never assigns, model.py:188-194; here it is populated so save/load round-trips)."""
        return dict(self._config)

    CONFIG_FILE = "config.json"           # file names of the reference (model.py:196-216), so either side's
    WEIGHTS_FILE = "pytorch_model.bin"    # checkpoints open with the other's loader
    _CTOR_FROM_CONFIG = {                 # config key -> (constructor argument, type)
        "num_rel": ("num_rel", int), "scorer_type": ("scorer_type", str), "gat_out_dim": ("gat_out_dim", int),
        "gat_heads": ("gat_heads", int), "dropout": ("dropout", float),
        "relation_attn_dropout": ("relation_attn_dropout", float), "gat_num_layers": ("gat_num_layers", int),
        "project_to_input_size": ("project_to_input_size", bool), "projection_layers": ("projection_layers", int),
        "projection_dropout": ("projection_dropout", float), "projection_hidden_dim": ("projection_hidden_dim", int),
    }

    def save_pretrained(self, output_dir: str, add_files: Optional[List[Tuple[str, Dict[str, Any]]]] = None) -> None:
        """Writes the weights, ``config.json`` and any extra (file name, JSON-able dict) pairs into ``output_dir``."""
        os.makedirs(output_dir, exist_ok=True)
        json_files = dict(add_files or [])
        json_files[self.CONFIG_FILE] = self.get_config()
        for file_name, payload in json_files.items():
            with open(os.path.join(output_dir, file_name), "w", encoding="utf-8") as fh:
                json.dump(payload, fh, ensure_ascii=False, indent=2)
        torch.save(self.state_dict(), os.path.join(output_dir, self.WEIGHTS_FILE))

    @staticmethod
    def load_from_pretrained(input_dir: str, *, node_emb: Optional[torch.Tensor] = None,
                             edge_index: Optional[torch.Tensor] = None, edge_type: Optional[torch.Tensor] = None,
                             map_location=None, precision: str = "fp32") -> "RelGATModel":
        """Rebuilds a model saved by ``save_pretrained``; the graph tensors are not part of a checkpoint and must
        be passed in (their feature width is checked against the saved configuration).  Returned in eval mode."""
        paths = {kind: os.path.join(input_dir, name) for kind, name in
                 (("Config", RelGATModel.CONFIG_FILE), ("Weights", RelGATModel.WEIGHTS_FILE))}
        for kind, path in paths.items():
            if not os.path.isfile(path):
                raise FileNotFoundError(f"{kind} file not found: {path}")
        with open(paths["Config"], "r", encoding="utf-8") as fh:
            saved = json.load(fh)
        width = None if node_emb is None else int(node_emb.size(1))
        if int(saved["input_dim"]) != width:
            raise ValueError(f"Input dim mismatch: config={saved['input_dim']} vs node_emb={width}")
        kwargs = {arg: cast(saved[key]) for key, (arg, cast) in RelGATModel._CTOR_FROM_CONFIG.items()}
        model = RelGATModel(node_emb, edge_index, edge_type, precision=precision, **kwargs)
        model.load_state_dict(torch.load(paths["Weights"], map_location=map_location), strict=True)
        return model.eval()
