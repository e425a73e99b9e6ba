#!/bin/bash
for m in 2 3 2 3 2; do FISR_PWC_UMMA=$m timeout 300 python tools/pwc_target.py 2>&1 | tail -1; done
