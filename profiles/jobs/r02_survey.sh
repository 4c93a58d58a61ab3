# round 2: ncu --set full over the common NON-baseline plans (looking for the kind of overhead the baseline kernels had:
# per-access predicates, spilled loop counters, un-prefetched loads); summaries only
mkdir -p gpurun_out
ncu --set full --clock-control none -k regex:'fast|strided|n13' -f -o /tmp/r02_survey python profiles/prof_plan.py c2dit r12 u12 s16 d18 t18 w12 n13 n13t n10 n14 n14t d18n13 > gpurun_out/r02_survey_ncu.log 2>&1
tail -2 gpurun_out/r02_survey_ncu.log
python profiles/summarize_ncu.py /tmp/r02_survey.ncu-rep gpurun_out/r02_survey.ncu_summary.txt > /dev/null 2>&1
ncu -i /tmp/r02_survey.ncu-rep --page source --csv --print-source sass > /tmp/r02_survey_src.csv 2>/dev/null
python profiles/ncu_source_hot.py /tmp/r02_survey_src.csv 10 > gpurun_out/r02_survey.ncu_hot.txt 2>&1
ls -la gpurun_out/r02_survey.*
