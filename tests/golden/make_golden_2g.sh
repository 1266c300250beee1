#!/bin/bash
# builds the reference 2 Gbp k=2 d=64 index + transforms + CPU search results with the UNMODIFIED reference tools
set -e
R=/root/repo/oracle/_ref
cd /tmp/fmdata/g2
$R/gfmiBaseLine_64bases_2step ref.fa 2000000000 > gfmi.log 2>&1
$R/tfmiAC_64bases_2step ref.fa.2000000000.64fmi2steps.fmi > tfmiac.log 2>&1
$R/tfmiBMP_64bases_2step ref.fa.2000000000.64fmi2steps.fmi > tfmibmp.log 2>&1
$R/fmIndexSearchCPU_64bases_2step ref.fa.2000000000.64fmi2steps.fmi reads.fa 100 1000000 > search_std.log 2>&1
$R/fmIndexSearchCPU_64bases_2step-ac ref.fa.2000000000.64fmi2steps.fmi.ac reads.fa 100 1000000 > search_ac.log 2>&1
md5sum ref.fa.2000000000.64fmi2steps.fmi* > md5.txt
echo DONE > done.flag
