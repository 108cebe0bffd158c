"""Summarise a --profile-csv file: per (kind, shape) launches, total/avg time, achieved GB/s and TFLOP/s."""
import collections, sys
KINDS = ['gemm_nt', 'gemm_tn', 'attention_fwd', 'attention_bwd', 'layernorm', 'lstm_gates', 'patch', 'other', 'conv']
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for line in open(sys.argv[1]):
    k, us, fl, by, d0, d1, d2 = line.strip().split(',')
    a = agg[(KINDS[int(k)], int(d0), int(d1), int(d2))]
    a[0] += 1; a[1] += float(us); a[2] += float(fl); a[3] += float(by)
tot = sum(a[1] for a in agg.values())
print(f'total {tot / 1e3:.2f} ms over {sum(a[0] for a in agg.values())} launches')
print(f'{"kind":14s} {"shape":>22s} {"n":>5s} {"tot ms":>8s} {"avg us":>8s} {"GB/s":>8s} {"TF/s":>7s} {"share":>6s}')
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    n, us, fl, by = a
    print(f'{key[0]:14s} {str(key[1:]):>22s} {n:5d} {us / 1e3:8.3f} {us / n:8.1f} {by / us / 1e3:8.0f} {fl / us / 1e6:7.1f} {us / tot:6.1%}')
