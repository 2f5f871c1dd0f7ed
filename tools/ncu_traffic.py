#!/usr/bin/env python
"""profiles/r2_traffic.json from the committed `ncu --set full` summary (tools/ncu_summary.py output).

    python tools/ncu_traffic.py profiles/r2_ncu_full_c3_x3.csv profiles/r2_traffic.json [C3]

DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per entry-point CALL: the mean over the captured launches of
each kernel, summed over the kernels one call launches.  bench.py reports it as `roofline.traffic`.
"""
import collections
import csv
import json
import sys

# kernel-name fragment -> entry point (bench.py's table names; fused variants are timed under the base name)
KERNEL_TO_ENTRY = [
    ('conv3x3_halo_k', 'lnst_conv3x3_bf16x3_tc'), ('conv3x3_tc_persist_k', 'lnst_gram_bwd_bf16x3_tc'),
    ('gram_split_tc_k', 'lnst_gram_diff_bf16x3_tc'), ('gram_finish_split3_k', 'lnst_gram_diff_bf16x3_tc'),
    ('gram_tc_k', 'lnst_gram_diff_bf16x3_tc'), ('gram_finish_split_k', 'lnst_gram_diff_bf16x3_tc'),
    ('avgpool2_split_bwd_k', 'lnst_avgpool2_bf16x3_bwd'), ('avgpool2_split_fwd_k', 'lnst_avgpool2_bf16x3_fwd'),
    ('conv_first_bwd_gray_col_k', 'lnst_conv_first_bwd_gray_x3_tc'), ('conv_first_fwd_gray_k', 'lnst_conv_first_fwd_gray_x3'),
    ('raymarch_rot_bwd_k', 'lnst_raymarch_bwd_box'), ('raymarch_fwd_tma_k', 'lnst_raymarch_fwd_tma'),
    ('smooth3_tma_k<1>', 'lnst_smooth3_relu_bwd_tma'), ('smooth3_tma_k<0>', 'lnst_smooth3_relu_fwd_tma'),
    ('splat_wavg_num3_k', 'lnst_splat_wavg_fwd_box'), ('splat_wavg_combine_box_k', 'lnst_splat_wavg_fwd_box'),
    ('splat_wavg_bwd3_k', 'lnst_splat_wavg_bwd_coef'), ('adam_iterate_dev', 'lnst_adam_iterate_dev'),
]


def main():
    src, dst = sys.argv[1], sys.argv[2]
    wl = sys.argv[3] if len(sys.argv) > 3 else 'C3'
    rows = list(csv.reader(open(src)))
    h = rows[0]
    ki = h.index('kernel')
    ri = [i for i, c in enumerate(h) if c.startswith('dram__bytes_read.sum')][0]
    wi = [i for i, c in enumerate(h) if c.startswith('dram__bytes_write.sum')][0]
    scale = {'Mbyte': 1e6, 'Kbyte': 1e3, 'Gbyte': 1e9, 'byte': 1.0}
    ru, wu = h[ri].split('[')[1].rstrip(']'), h[wi].split('[')[1].rstrip(']')
    per_kernel = collections.OrderedDict()
    for r in rows[1:]:
        frag = next((f for f, _ in KERNEL_TO_ENTRY if f in r[ki]), None)
        if frag is None:
            continue
        per_kernel.setdefault(frag, []).append(float(r[ri]) * scale[ru] + float(r[wi]) * scale[wu])
    entry = collections.OrderedDict()
    for frag, e in KERNEL_TO_ENTRY:
        if frag in per_kernel:
            v = per_kernel[frag]
            entry[e] = entry.get(e, 0.0) + sum(v) / len(v)
    out = {}
    try:
        out = json.load(open(dst))
    except Exception:
        pass
    out[wl] = {k: int(v) for k, v in entry.items()}
    out[wl]['_source'] = ('%s: dram__bytes_read.sum + dram__bytes_write.sum per entry-point call (mean over the launches '
                          'captured in one bf16x3 step, ncu --set full --clock-control none); tools/ncu_traffic.py' % src)
    json.dump(out, open(dst, 'w'), indent=1)
    print(json.dumps(out[wl], indent=1))


if __name__ == '__main__':
    main()
