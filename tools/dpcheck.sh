T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
for v in "A=1" "LIVAE_UPFOLD=0" "LIVAE_HALO16=1" "LIVAE_UPFOLD=0 LIVAE_HALO16=1"; do
  echo "== $v"; env $v timeout 300 $T bench.py --gpus 2 --check-dp 2>/dev/null | grep -o '"grad_rel_l2_after_clip": [^,]*, "param_max_abs_diff_after_adamw": [^,]*'
done
