#!/bin/bash
# Extracts the SASS/PTX evidence the profiling recipe asks for (B200_PROFILING.md,
# "What proves a Blackwell-native kernel") from the built extension into profiles/.
set -eu
SO=mpi4torch_b200/_lib/_m4t_C.so
OUT=profiles/sass
mkdir -p $OUT
cuobjdump -sass $SO > /tmp/m4t_sass.txt
python - <<'PY'
import re, collections
txt = open('/tmp/m4t_sass.txt').read()
funcs = re.split(r'\n\s*Function : ', txt)
want = {
  'gemm_bf16_tn_2cta_kernelILb1ELi1ELb0E': 'fused_allreduce_gemm_mse_cta_pair',
  'gemm_bf16_tn_2cta_kernelILb0ELi0ELb0E': 'gemm_tcgen05_cta_pair',
  'gemm_bf16_tn_kernelILb1ELi1E': 'fused_allreduce_gemm_mse_single_cta',
  'gemm_bf16_tn_kernelILb0ELi0E': 'gemm_tcgen05_single_cta',
  'allreduce_pipelined_kernelILNS_5DTypeE7ELNS_8ReduceOpE2ELNS_8NvlsKindE2E': 'allreduce_nvls_pipelined_bf16_sum',
  'allreduce_twoshot_kernelILNS_5DTypeE7ELNS_8ReduceOpE2ELNS_8NvlsKindE2E': 'allreduce_nvls_bf16_sum',
  'allreduce_oneshot_kernelILNS_5DTypeE7ELNS_8ReduceOpE2E': 'allreduce_oneshot_bf16_sum',
  'wgrad_bf16_nt_2cta_kernelILb1E': 'fused_wgrad_reduce_scatter_sgd_cta_pair',
  'wgrad_bf16_nt_2cta_kernelILb0E': 'wgrad_mn_major_cta_pair',
  'gemm_bf16_tn_2cta_kernelILb0ELi1ELb0E': 'gemm_mse_epilogue_cta_pair',
  'gemm_bf16_tn_2cta_kernelILb0ELi0ELb1E': 'gemm_nn_dgrad_cta_pair',
  'p2p_flag_wait_kernel': 'p2p_copy_engine_flag_wait',
  'slab_reduce_vec_kernelILNS_5DTypeE7ELNS_8ReduceOpE2ELNS_8NvlsKindE2E': 'reduce_scatter_nvls_bf16_sum',
  'p2p_recv_kernel': 'p2p_recv',
  'slab_pull_kernelILi16E': 'slab_pull_16B',
  'bcast_kernel': 'bcast',
}
pat = re.compile(r'\b(UTCHMMA[.\w]*|UTCQMMA|UTMALDG[.\w]*|UTMASTG[.\w]*|LDTM[.\w]*|STTM[.\w]*|UTCBAR[.\w]*|LDGMC[.\w]*|STG\.[\w.]*MC[\w.]*|REDG?[.\w]*MC[.\w]*|MULTIMEM[.\w]*|LDG\.E\.128[.\w]*|LDG\.E\.NA\.128[.\w]*|STG\.E\.128[.\w]*|SYNCS[.\w]*|HMMA[.\w]*|MEMBAR[.\w]*|ST\.E[.\w]*STRONG\.SYS|LD\.E[.\w]*STRONG\.SYS)\b')
summary = []
seen = set()
for f in funcs[1:]:
    name = f.split('\n', 1)[0].strip()
    for key, label in want.items():
        if key in name and label not in seen:
            seen.add(label)
            ops = collections.Counter(m.group(1) for m in pat.finditer(f))
            open(f'profiles/sass/{label}.sass', 'w').write('Function : ' + f)
            summary.append((label, name, ops))
            break
with open('profiles/sass/SUMMARY.md', 'w') as out:
    out.write('# SASS evidence (cuobjdump -sass of mpi4torch_b200/_lib/_m4t_C.so, sm_100a)\n\n')
    out.write('PTX -> SASS: tcgen05.mma = UTCHMMA, tcgen05.ld = LDTM, TMA = UTMALDG, multimem.ld_reduce = LDGMC, '
              'multimem.st / multimem.red = STG/RED with the .MC* modifiers, mbarrier = SYNCS.\n\n')
    for label, name, ops in summary:
        out.write(f'## {label}\n`{name[:160]}`\n\n')
        for op, n in sorted(ops.items(), key=lambda kv: -kv[1]):
            out.write(f'- `{op}` x {n}\n')
        out.write('\n')
print(open('profiles/sass/SUMMARY.md').read()[:3000])
PY
