"""Helper (not a test; run as a sub-process with PYTHONPATH = [shims, reference checkout, repo root]): glue.model_forward against the
reference's OWN GATModel.forward / forward_and_return_attention / calc_attention_norm (models/GATModel.py:118-234) on the CPU.

model_forward drives layers through the B200 layer's glue contract (forward(..., skip=), output_activation, output_dropout,
attention_norm / attention_norm_value).  Here that contract is implemented by a thin wrapper around the REFERENCE's GATLayer with
plain torch ops, so what is compared is exactly model_forward's sequencing -- which layer gets which switch, where the skip rows
come from, what is returned -- against the reference's code, on the PPI, PATTERN and Cora stacks.  Prints MODEL_FORWARD_OK."""
import sys
import types

import torch
import torch.nn.functional as F

from models.GATModel import GATModel            # the reference's file (through the stand-ins for Lightning / PyG)
from models.utils import sum_over_neighbourhood
from run_config import LayerType
from gat_pytorch_b200.glue import model_forward


class GlueContract(torch.nn.Module):
    """The opt-in glue of gat_pytorch_b200.GATLayer (include/gat_b200.h "OUTPUT GLUE"), restated with torch ops around the
    reference layer: y = dropout(E(layer(x) + skip)); attention_norm_value = sum |alpha*deg - 1| / E' (GATModel.py:196-224)."""

    def __init__(self, ref_layer):
        super().__init__()
        self.ref = ref_layer
        self.output_activation, self.output_dropout, self.attention_norm, self.attention_norm_value = None, 0.0, False, None

    def forward(self, x, edge_index, return_attention_weights=False, skip=None):
        out, (ei2, alpha) = self.ref(x, edge_index, return_attention_weights=True)
        if skip is not None:
            out = out + skip
        if self.output_activation == "elu":
            out = F.elu(out)
        if self.training and self.output_dropout > 0:
            out = F.dropout(out, p=self.output_dropout, training=True)
        if self.attention_norm:
            deg = sum_over_neighbourhood(torch.ones_like(alpha[:, 0]), ei2[1], alpha[:, 0].size(), broadcast_back=True)
            self.attention_norm_value = torch.norm(alpha * deg[:, None] - 1.0, p=1) / ei2.size(1)
        return (out, (ei2, alpha)) if return_attention_weights else out


def build(cfg):
    torch.manual_seed(7)
    m = GATModel(layer_type=LayerType.GATLayer, dataset="x", num_classes=cfg["out"], num_input_node_features=cfg["f_in"],
                 num_layers=len(cfg["heads"]), num_heads_per_layer=cfg["heads"], heads_concat_per_layer=cfg["concat"],
                 head_output_features_per_layer=[cfg["f_in"]] + cfg["feats"], add_skip_connection=cfg["skip"], dropout=0.0,
                 l2_reg=0.0, learning_rate=0.01, batch_size=1, num_epochs=1, const_attention=False)
    return m


def main():
    cfgs = {"ppi": dict(f_in=10, out=5, heads=[4, 4, 6], feats=[8, 8, 5], concat=[True, True, False], skip=[False, True, False]),
            "pattern": dict(f_in=3, out=1, heads=[4, 4, 4, 1], feats=[3, 6, 3, 1], concat=[True, True, True, False], skip=[True] * 4),
            "cora": dict(f_in=12, out=7, heads=[8, 1], feats=[8, 7], concat=[True, False], skip=[False, False])}
    g = torch.Generator().manual_seed(3)
    n = 40
    ei = torch.randint(0, n, (2, 300), generator=g)
    for name, cfg in cfgs.items():
        model = build(cfg)
        x = torch.randn((n, cfg["f_in"]), generator=g)
        data = types.SimpleNamespace(x=x, edge_index=ei)
        wrapped = types.SimpleNamespace(gat_layer_list=torch.nn.ModuleList([GlueContract(l) for l in model.gat_layer_list]),
                                        skip_layer_list=model.skip_layer_list, add_skip_connection=model.add_skip_connection,
                                        heads_concat_per_layer=model.heads_concat_per_layer, num_heads_per_layer=model.num_heads_per_layer,
                                        head_output_features_per_layer=model.head_output_features_per_layer, dropout=model.dropout,
                                        training=False)
        for training in (False, True):          # dropout 0: both modes are deterministic
            model.train(training)
            wrapped.training = training
            wrapped.gat_layer_list.train(training)
            want = model.forward(data)
            got = model_forward(wrapped, data)
            assert torch.allclose(got, want, rtol=1e-6, atol=1e-7), (name, "forward", float((got - want).abs().max()))
            want_x, want_ei, want_att = model.forward_and_return_attention(data)
            got_x, got_ei, got_att = model_forward(wrapped, data, True)
            assert torch.allclose(got_x, want_x, rtol=1e-6, atol=1e-7) and torch.equal(got_ei, want_ei), (name, "with attention")
            assert len(got_att) == len(want_att) and all(torch.allclose(a, b, rtol=1e-6, atol=1e-8) for a, b in zip(got_att, want_att))
            model.logger = None                 # calc_attention_norm writes histograms when a logger is attached (GATModel.py:214)
            want_norm = model.calc_attention_norm(want_ei, want_att)
            got_out, got_norm = model_forward(wrapped, data, attention_norm=True)
            assert torch.allclose(got_out, want, rtol=1e-6, atol=1e-7)
            assert abs(float(got_norm) - float(want_norm)) <= 1e-6 * abs(float(want_norm)) + 1e-9, (name, float(got_norm), float(want_norm))
        assert all(l.output_activation is None and l.output_dropout == 0.0 and l.attention_norm is False for l in wrapped.gat_layer_list)
    print("MODEL_FORWARD_OK", flush=True)


if __name__ == "__main__":
    main()
