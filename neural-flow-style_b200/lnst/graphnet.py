"""Loss network given as a frozen TensorFlow GraphDef -- the reference's inception5h path
(``styler_base.py:17-31``: parse ``tensorflow_inception_graph.pb``; ``:53-57``: ``tf.import_graph_def(graph_def,
{'input': vgg.preprocess(d)})``; ``:91-94``: layers are read by tensor name ``import/<layer>:0``).

The graph is interpreted, not hard-coded: nodes are parsed by ``lnst.graphdef`` and each TF op maps to one C-ABI
entry point of ``csrc/graphnet.cu`` (fp32 NHWC):

    Conv2D (+ the BiasAdd that consumes it)  lnst_conv2d_f32 / lnst_conv2d_bwd_data_f32
    Relu                                      lnst_relu_fwd / lnst_relu_bwd
    MaxPool                                   lnst_maxpool_fwd / lnst_maxpool_bwd
    LRN                                       lnst_lrn_fwd / lnst_lrn_bwd
    Concat / ConcatV2 (channel axis)          lnst_copy_channels
    AvgPool                                   lnst_avgpool_fwd / lnst_avgpool_bwd     (the head's avgpool0)
    Reshape to [-1, C], MatMul                a view / lnst_conv2d_f32 as a 1x1 convolution (softmax2_pre_activation:
                                              the class logits the reference's top_k content target reads,
                                              styler_base.py:240-246); 2-D tensors are carried as [n, rows, 1, C]
    Identity, Placeholder, Const              --

Fusions decided per call from the requested layers (``_fusion``): Conv2D + BiasAdd always; + Relu when the pre-ReLU
tensor is not itself a requested layer (the ReLU gradient is then folded into the operand load of the data-gradient
GEMM); branch convolutions of an inception module write straight into their channel slice of the module's concat buffer
and read their cotangent slice in place (no concat / slice copies).

Only the sub-graph between ``input`` and the requested layers runs.  Weights are frozen: the backward pass
computes data gradients only, accumulating where a tensor feeds several consumers.  ``GraphNet`` offers the same
methods as ``lnst.vgg.LossNet`` (``forward / backward / gram / gram_grad / content ...``) so ``StylerBase`` drives
either network.  The arithmetic is fp32 on CUDA cores (``conv_math`` is ignored): the tcgen05 path covers the
3x3 VGG stack only.
"""
import numpy as np
import torch

from . import graphdef, ops

f32 = torch.float32
_SUPPORTED = ('Placeholder', 'Const', 'Conv2D', 'BiasAdd', 'Relu', 'MaxPool', 'AvgPool', 'LRN', 'Concat', 'ConcatV2',
              'Identity', 'Reshape', 'MatMul')


def _clean(name):
    """'import/x:0', 'x:0', '^x' -> 'x' (tensor names as the reference spells them, styler_base.py:94)"""
    name = name[7:] if name.startswith('import/') else name
    name = name[1:] if name.startswith('^') else name
    return name.split(':')[0]


class GraphNet(object):
    def __init__(self, nodes, device, pool1=False, input_name='input'):
        self.device = torch.device(device)
        self.math = 'fp32'
        self.model = 'graphdef'
        self.input_name = input_name
        nodes = [graphdef.Node(n.name, n.op, n.inputs, n.attr) for n in nodes]      # private copy (pool1 edits attrs)
        self.nodes = {n.name: n for n in nodes}
        self.order = [n.name for n in nodes]                   # GraphDefs are stored in topological order
        if input_name not in self.nodes:
            raise KeyError('graph has no %r node' % input_name)
        self.const = {}
        for n in nodes:
            if n.op == 'Const' and n.attr.get('value') is not None:
                self.const[n.name] = n.attr['value']
        if pool1:                                              # styler_base.py:26-31: stride-1 first convolution
            for n in nodes:
                if 'conv2d0_pre_relu/conv' in n.name:
                    s = list(n.attr['strides'])
                    s[1:3] = [1, 1]
                    n.attr['strides'] = s
        self._dev_w = {}
        self._plan_cache, self._fusion_cache = {}, {}            # per requested-layer set (host-side graph walks)
        # Conv2D whose only consumer is a BiasAdd: run as one kernel (the un-biased tensor is not materialised)
        consumers = {}
        for n in nodes:
            for i in n.inputs:
                consumers.setdefault(_clean(i), []).append(n.name)
        self.consumers = consumers
        self.fused_bias = {}
        for n in nodes:
            if n.op == 'BiasAdd':
                src = self.nodes.get(_clean(n.inputs[0]))
                if src is not None and src.op == 'Conv2D' and consumers.get(src.name) == [n.name]:
                    self.fused_bias[n.name] = src.name

    @classmethod
    def from_file(cls, path, device, pool1=False):
        return cls(graphdef.load(path), device, pool1=pool1)

    # ---- LossNet surface -----------------------------------------------------------------------------
    def gray_path(self):
        return False

    def relu_masked(self, name):
        """Gradients are held w.r.t. the named tensor itself (Relu is its own node)."""
        return False

    def feature_pixels(self, acts, name):
        f = acts[_clean(name)]
        return f.shape[1] * f.shape[2]

    def _w(self, name):
        t = self._dev_w.get(name)
        if t is None:
            t = torch.tensor(np.ascontiguousarray(self.const[name], np.float32)).to(self.device)
            self._dev_w[name] = t
        return t

    def _plan(self, wanted):
        """Nodes (file order) on a path from the input to a wanted tensor; Const inputs are not activations."""
        key = tuple(sorted(_clean(w) for w in wanted))
        if key not in self._plan_cache:
            self._plan_cache[key] = self._plan_walk(wanted)
        return self._plan_cache[key]

    def _plan_walk(self, wanted):
        need, stack = set(), [_clean(w) for w in wanted if 'input' != _clean(w)]
        while stack:
            name = stack.pop()
            if name in need or name in self.const or name == self.input_name:
                continue
            node = self.nodes.get(name)
            if node is None:
                raise KeyError('%s is not a tensor of the loss network graph' % name)
            if node.op not in _SUPPORTED:
                raise NotImplementedError('op %s (%s) is not built for GraphDef loss networks' % (node.op, name))
            need.add(name)
            stack += [_clean(i) for i in node.inputs]
        return [n for n in self.order if n in need]

    def _data_inputs(self, node):
        ins = [_clean(i) for i in node.inputs if not i.startswith('^')]
        if node.op == 'Concat':
            return ins[1:]                                     # concat_dim first
        if node.op == 'ConcatV2':
            return ins[:-1]                                    # axis last
        if node.op in ('Conv2D', 'BiasAdd', 'Reshape', 'MatMul'):
            return ins[:1]
        return ins

    def _concat_axis(self, node):
        ins = [_clean(i) for i in node.inputs]
        axis = int(np.asarray(self.const[ins[0] if node.op == 'Concat' else ins[-1]]).reshape(-1)[0])
        if axis not in (3, -1):
            raise NotImplementedError('%s: only channel-axis concatenation is built' % node.name)

    def _conv_attrs(self, node):
        strides = list(node.attr.get('strides', [1, 1, 1, 1]))
        if strides[1] != strides[2] or strides[0] != 1 or strides[3] != 1:
            raise NotImplementedError('%s: strides %s' % (node.name, strides))
        fmt = node.attr.get('data_format', b'NHWC')
        if fmt not in (b'NHWC', None):
            raise NotImplementedError('%s: data_format %s' % (node.name, fmt))
        return int(strides[1]), (node.attr.get('padding') or b'SAME').decode()

    def _fusion(self, wanted):
        """Per call: which Relu nodes run inside their convolution, and which of those write straight into a concat.

        relu_of[R] = B      Relu R whose input is a Conv2D+BiasAdd B that nobody else reads and that is not itself a
                            requested layer (a content layer such as ``mixed4d_3x3_bottleneck_pre_relu`` is, and keeps
                            the separate Relu pass): one kernel, relu = 1.
        slot[R] = (C, off)  such an R whose only consumer is the channel-axis Concat C, every input of C qualifying
                            and none of them requested: R's convolution writes channels [off, off+ch) of C's buffer,
                            the concat copies disappear, the backward pass reads the cotangent slice in place.
        """
        key = tuple(sorted(_clean(w) for w in wanted))
        if key in self._fusion_cache:
            return self._fusion_cache[key]
        want = {_clean(w) for w in wanted}
        relu_of = {}
        for name, node in self.nodes.items():
            if node.op != 'Relu':
                continue
            b = _clean(node.inputs[0])
            if b in self.fused_bias and b not in want and self.consumers.get(b) == [name]:
                relu_of[name] = b
        slot, width = {}, {}
        for name, node in self.nodes.items():
            if node.op not in ('Concat', 'ConcatV2'):
                continue
            ins = self._data_inputs(node)
            if not ins or any(i not in relu_of or i in want or self.consumers.get(i) != [name] for i in ins):
                continue
            off = 0
            for i in ins:
                conv = self.nodes[self.fused_bias[relu_of[i]]]
                slot[i] = (name, off)
                off += int(self.const[_clean(conv.inputs[1])].shape[-1])
            width[name] = off
        self._fusion_cache[key] = (relu_of, slot, width)
        return relu_of, slot, width

    def forward(self, x, wanted, gray=None):
        """x [n,H,W,3] mean-subtracted net input.  Returns {tensor name: fp32 [n,h,w,C]} for every materialised node
        (plus ``'__fusion__'``, the plan the backward pass must follow)."""
        relu_of, slot, width = self._fusion(wanted)
        fused_pre = set(relu_of.values())
        acts = {self.input_name: x, '__fusion__': (relu_of, slot)}
        for name in self._plan(wanted):
            node = self.nodes[name]
            ins = self._data_inputs(node)
            if node.op == 'Conv2D':
                if name in self.fused_bias.values():
                    continue                                   # produced together with its BiasAdd
                stride, padding = self._conv_attrs(node)
                acts[name] = ops.conv2d_f32(acts[ins[0]], self._w(_clean(node.inputs[1])), None, stride, padding)
            elif node.op == 'BiasAdd':
                if name in fused_pre:
                    continue                                   # produced by its Relu (conv + bias + relu in one kernel)
                if name in self.fused_bias:
                    conv = self.nodes[self.fused_bias[name]]
                    stride, padding = self._conv_attrs(conv)
                    acts[name] = ops.conv2d_f32(acts[_clean(conv.inputs[0])], self._w(_clean(conv.inputs[1])),
                                                self._w(_clean(node.inputs[1])), stride, padding)
                else:
                    acts[name] = acts[ins[0]] + self._w(_clean(node.inputs[1]))
            elif node.op == 'Relu':
                if name in relu_of:
                    bias_node = self.nodes[relu_of[name]]
                    conv = self.nodes[self.fused_bias[relu_of[name]]]
                    stride, padding = self._conv_attrs(conv)
                    src, w = acts[_clean(conv.inputs[0])], self._w(_clean(conv.inputs[1]))
                    if name in slot:                           # straight into the concat buffer
                        cat, off = slot[name]
                        if cat not in acts:
                            OH, OW, _, _ = ops.conv_out(src.shape, w.shape, stride, padding)
                            acts[cat] = torch.empty(src.shape[0], OH, OW, width[cat], dtype=f32, device=self.device)
                        ops.conv2d_f32(src, w, self._w(_clean(bias_node.inputs[1])), stride, padding, relu=True,
                                       out=acts[cat], ch_off=off)
                    else:
                        acts[name] = ops.conv2d_f32(src, w, self._w(_clean(bias_node.inputs[1])), stride, padding,
                                                    relu=True)
                else:
                    acts[name] = ops.relu_fwd(acts[ins[0]])
            elif node.op == 'MaxPool':
                k, stride, padding = self._pool_attrs(node)
                acts[name] = ops.maxpool_fwd(acts[ins[0]], k, stride, padding)
            elif node.op == 'LRN':
                r, bias, alpha, beta = self._lrn_attrs(node)
                acts[name] = ops.lrn_fwd(acts[ins[0]], r, bias, alpha, beta)
            elif node.op == 'AvgPool':
                k, stride, padding = self._pool_attrs(node)
                acts[name] = ops.avgpool_fwd(acts[ins[0]], k, stride, padding)
            elif node.op == 'Reshape':                         # [n,h,w,C] -> [-1, C], carried as [n, h*w, 1, C]
                src = acts[ins[0]]
                shp = [int(v) for v in np.asarray(self.const[_clean(node.inputs[1])]).reshape(-1)]
                if len(shp) != 2 or shp[0] != -1 or shp[1] != src.shape[-1]:
                    raise NotImplementedError('%s: only Reshape to [-1, channels] is built (got %s)' % (name, shp))
                acts[name] = src.reshape(src.shape[0], -1, 1, src.shape[-1])
            elif node.op == 'MatMul':                          # rows x [K, N]: a 1x1 convolution
                if node.attr.get('transpose_a') or node.attr.get('transpose_b'):
                    raise NotImplementedError('%s: transposed MatMul' % name)
                w = self._w(_clean(node.inputs[1]))
                acts[name] = ops.conv2d_f32(acts[ins[0]], w.reshape(1, 1, w.shape[0], w.shape[1]), None, 1, 'SAME')
            elif node.op in ('Concat', 'ConcatV2'):
                self._concat_axis(node)
                if name in width:
                    continue                                   # its inputs wrote acts[name] directly
                parts = [acts[i] for i in ins]
                out = torch.empty(parts[0].shape[:3] + (sum(p.shape[-1] for p in parts),), dtype=f32, device=self.device)
                off = 0
                for p in parts:
                    ops.copy_channels(p, 0, out, off, p.shape[-1])
                    off += p.shape[-1]
                acts[name] = out
            elif node.op == 'Identity':
                acts[name] = acts[ins[0]]
            elif node.op == 'Placeholder':
                raise KeyError('placeholder %s is not fed (only %r is)' % (name, self.input_name))
        return acts

    def _pool_attrs(self, node):
        ks, st = list(node.attr['ksize']), list(node.attr['strides'])
        if ks[1] != ks[2] or st[1] != st[2]:
            raise NotImplementedError('%s: ksize %s strides %s' % (node.name, ks, st))
        return int(ks[1]), int(st[1]), (node.attr.get('padding') or b'SAME').decode()

    @staticmethod
    def _lrn_attrs(node):
        a = node.attr                                          # op defaults of tf.nn.lrn
        return (int(a.get('depth_radius', 5)), float(a.get('bias', 1.0)), float(a.get('alpha', 1.0)),
                float(a.get('beta', 0.5)))

    def backward(self, x, acts, wanted, add_loss_grad, loss_layers, gray=False):
        """d loss / d x [n,H,W,3].  ``add_loss_grad(name, g)`` adds the loss terms living on tensor ``name`` into g
        (None = nothing accumulated yet) and returns the buffer."""
        plan = self._plan(wanted)
        relu_of, slot = acts['__fusion__']
        fused_pre = set(relu_of.values())
        loss_names = {_clean(l): l for l in loss_layers}
        grads = {}
        sliced = {}                                            # R -> (cotangent of its concat, channel offset)

        def acc(name, shape_like):
            """(buffer, accumulate flag) for the gradient of tensor ``name``"""
            if name in grads:
                return grads[name], True
            grads[name] = torch.empty_like(shape_like)
            return grads[name], False

        for name in reversed(plan):
            node = self.nodes[name]
            if name in fused_pre or (node.op == 'Conv2D' and name in self.fused_bias.values()):
                continue                                       # handled at the node that produced the tensor
            if name in loss_names:
                grads[name] = add_loss_grad(loss_names[name], grads.get(name))
            g = grads.pop(name, None)
            ins = self._data_inputs(node)
            if node.op == 'Relu' and name in relu_of:          # conv + bias + relu: mask folded into the dgrad load
                conv = self.nodes[self.fused_bias[relu_of[name]]]
                src = _clean(conv.inputs[0])
                stride, padding = self._conv_attrs(conv)
                if name in sliced:
                    gcat, off = sliced.pop(name)
                    gx, a = acc(src, acts[src])
                    ops.conv2d_bwd_data_f32(gcat, self._w(_clean(conv.inputs[1])), acts[src].shape, stride, padding, gx,
                                            a, ch_off=off, relu_y=acts[slot[name][0]])
                elif g is not None:
                    gx, a = acc(src, acts[src])
                    ops.conv2d_bwd_data_f32(g, self._w(_clean(conv.inputs[1])), acts[src].shape, stride, padding, gx, a,
                                            relu_y=acts[name])
                continue
            if g is None:
                continue
            if node.op == 'Conv2D' or name in self.fused_bias:
                conv = self.nodes[self.fused_bias[name]] if name in self.fused_bias else node
                src = _clean(conv.inputs[0])
                stride, padding = self._conv_attrs(conv)
                gx, a = acc(src, acts[src])
                ops.conv2d_bwd_data_f32(g, self._w(_clean(conv.inputs[1])), acts[src].shape, stride, padding, gx, a)
            elif node.op in ('BiasAdd', 'Identity'):
                gx, a = acc(ins[0], acts[ins[0]])
                ops.copy_channels(g, 0, gx, 0, g.shape[-1], accumulate=a)
            elif node.op == 'Relu':
                gx, a = acc(ins[0], acts[ins[0]])
                ops.relu_bwd(g, acts[name], gx, a)
            elif node.op == 'MaxPool':
                k, stride, padding = self._pool_attrs(node)
                gx, a = acc(ins[0], acts[ins[0]])
                ops.maxpool_bwd(g, acts[ins[0]], k, stride, padding, gx, a)
            elif node.op == 'LRN':
                r, bias, alpha, beta = self._lrn_attrs(node)
                gx, a = acc(ins[0], acts[ins[0]])
                ops.lrn_bwd(g, acts[ins[0]], r, bias, alpha, beta, gx, a)
            elif node.op == 'AvgPool':
                k, stride, padding = self._pool_attrs(node)
                gx, a = acc(ins[0], acts[ins[0]])
                ops.avgpool_bwd(g, acts[ins[0]].shape, k, stride, padding, gx, a)
            elif node.op == 'Reshape':
                gx, a = acc(ins[0], acts[ins[0]])
                ops.copy_channels(g, 0, gx, 0, g.shape[-1], accumulate=a)      # same memory order, other shape
            elif node.op == 'MatMul':
                w = self._w(_clean(node.inputs[1]))
                gx, a = acc(ins[0], acts[ins[0]])
                ops.conv2d_bwd_data_f32(g, w.reshape(1, 1, w.shape[0], w.shape[1]), acts[ins[0]].shape, 1, 'SAME', gx, a)
            elif node.op in ('Concat', 'ConcatV2'):
                if all(i in slot and slot[i][0] == name for i in ins):
                    for i in ins:                              # the branches read their slice of g in place
                        sliced[i] = (g, slot[i][1])
                    continue
                off = 0
                for i in ins:
                    ch = acts[i].shape[-1]
                    gx, a = acc(i, acts[i])
                    ops.copy_channels(g, off, gx, 0, ch, accumulate=a)
                    off += ch
        return grads.get(self.input_name)

    # ---- losses on end points: same kernels as the fp32 VGG path ---------------------------------------
    def features_f32(self, acts, name):
        return acts[_clean(name)]

    def gram(self, acts, name, Gs, weight, loss, mask=None):
        if mask is not None:
            raise NotImplementedError('style_mask with a GraphDef loss network')
        f = acts[_clean(name)]
        n, P, ch = f.shape[0], f.shape[1] * f.shape[2], f.shape[3]
        out = {'G': [], 'den': [], 'weight': weight}
        for v in range(n):
            G = torch.empty(ch, ch, dtype=f32, device=self.device)
            den = 2.0 * P * ch
            ops.gram_diff(f[v].reshape(P, ch), den, Gs, weight, G, loss[v:v + 1] if loss is not None else None)
            out['G'].append(G)
            out['den'].append(den)
        return out

    def gram_values(self, handle):
        return torch.stack(handle['G'], 0)

    def gram_grad(self, acts, name, handle, coef, g, relu_mask):
        f = acts[_clean(name)]
        n, P, ch = f.shape[0], f.shape[1] * f.shape[2], f.shape[3]
        beta = 1.0
        if g is None:
            g, beta = torch.empty_like(f), 0.0
        for v in range(n):
            ops.gram_bwd(f[v].reshape(P, ch), handle['G'][v], 4.0 * handle['weight'] / handle['den'][v], beta, 0,
                         g[v].reshape(P, ch))
        return g

    def content(self, acts, name, channel, weight, loss, g, relu_mask, target=None, amp=1.0):
        f = acts[_clean(name)]
        n, P, ch = f.shape[0], f.shape[1] * f.shape[2], f.shape[3]
        beta = 1.0
        if g is None:
            g, beta = torch.empty_like(f), 0.0
        for v in range(n):
            if target is not None:
                ops.content_mse(f[v], target, amp, weight, loss[v:v + 1], g[v], beta, 0)
            else:
                ops.content_loss(f[v].reshape(P, ch), channel, weight, loss[v:v + 1], g[v].reshape(P, ch), beta, 0)
        return g
