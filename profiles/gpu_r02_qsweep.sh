out=gpurun_out/r02e_qsweep.jsonl; : > $out
for v in 0 1 2; do
  timeout 120 python profiles/pass_times.py --cfg cfg3 --all-chains 2 --var-h $v --var-v $v >> $out 2>> ${out}.err
done
timeout 120 python profiles/pass_times.py --cfg cfg3 >> $out 2>> ${out}.err
timeout 120 python profiles/pass_times.py --cfg u8kdil --all-chains 2 >> $out 2>> ${out}.err
timeout 120 python profiles/pass_times.py --cfg u8kdil >> $out 2>> ${out}.err
cut -c1-300 $out; tail -3 ${out}.err
