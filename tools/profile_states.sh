# ncu --set full of ONE full-grid launch of the dense fused kernel in the steady / cold state of a Gauss-Newton run (tools/profile_dense.py), summaries + per-source-line view.
mkdir -p gpurun_out
SYM='_ZN3pvb11k_associateILi10ELb1ELi6ELb0ELi4ELb1EEEvNS_9AssocArgsE'
for st in steady cold; do
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_associate -f -o /tmp/r2k_$st python tools/profile_dense.py $st 2>&1 | grep -E "kernel_ms|Error|error" 
  python profiles/summarize.py /tmp/r2k_$st.ncu-rep > gpurun_out/r2k_$st.txt 2>&1
  python tools/ncu_lines.py /tmp/r2k_$st.ncu-rep "$SYM" 60 > gpurun_out/r2k_${st}_lines.txt 2>&1
done
