#!/bin/bash
# final tree in one call: full GPU suite, headline bench, reference arm, other configs, then the evidence captures
bash scripts/gpu_job_final.sh
bash scripts/gpu_job_evidence.sh > gpurun_out/evidence_stdout.log 2>&1
tail -n 30 gpurun_out/evidence_stdout.log
