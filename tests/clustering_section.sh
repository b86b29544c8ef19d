#!/bin/bash
# The call sequence of the clustering section of the reference driver (src/pacybara.sh:425-489),
# reduced to the commands that touch the hot path, resolved by bare name on PATH exactly as the
# reference does -- so putting bin/shims and bin first on PATH swaps the implementation in.
#   usage: clustering_section.sh <EXTRACTDIR> <CLUSTERDIR> <MAXDIFF> <MINQUAL> <MINJACCARD> <MINMATCHES> <uptag|virtual>
set -euo pipefail
EXTRACTDIR=$1; CLUSTERDIR=$2; MAXDIFF=$3; MINQUAL=$4; MINJACCARD=$5; MINMATCHES=$6; CLUSTERMODE=$7
mkdir -p "$CLUSTERDIR/qc"
if [[ $CLUSTERMODE == "virtual" ]]; then BC=bcExtract_combo.fastq.gz; else BC=bcExtract_1.fastq.gz; fi
zcat "$EXTRACTDIR/$BC" | pacybara_precluster.py | gzip -c > "$CLUSTERDIR/bcPreclust.fastq.gz"
zcat "$CLUSTERDIR/bcPreclust.fastq.gz" | grep size | cut -f2,2 -d= | sort -n | uniq -c > "$CLUSTERDIR/qc/bcPreclust_distr.txt"
seqret -sequence <(zcat "$CLUSTERDIR/bcPreclust.fastq.gz") -outseq "$CLUSTERDIR/bcPreclust.fasta"
mkdir -p "$CLUSTERDIR/db"
bowtie2-build "$CLUSTERDIR/bcPreclust.fasta" "$CLUSTERDIR/db/bcDB"
rm "$CLUSTERDIR/bcPreclust.fasta"
bowtie2 --no-head --norc --very-sensitive --all -x "$CLUSTERDIR/db/bcDB" -U "$CLUSTERDIR/bcPreclust.fastq.gz" 2>/dev/null \
  | awk '{if($1!=$3){print}}' | gzip -c > "$CLUSTERDIR/bcMatches.sam.gz"
rm -r "$CLUSTERDIR/db"
pacybara_calcEdits.R "$CLUSTERDIR/bcMatches.sam.gz" "$CLUSTERDIR/bcPreclust.fastq.gz" --maxErr "$((MAXDIFF*2))" \
  --output "$CLUSTERDIR/editDistance.csv.gz"
zcat "$CLUSTERDIR/editDistance.csv.gz" | tail -n +2 | cut -f5,5 -d, | sort -n | uniq -c > "$CLUSTERDIR/qc/edDistr.txt"
python3 -u "$(command -v pacybara_cluster.py)" --uptagBarcodeFile "$EXTRACTDIR/bcExtract_1.fastq.gz" \
  --out "$CLUSTERDIR/clusters.csv.gz" --minJaccard "$MINJACCARD" --minMatches "$MINMATCHES" --maxDiff "$MAXDIFF" \
  --minQual "$MINQUAL" "$CLUSTERDIR/editDistance.csv.gz" "$EXTRACTDIR/genoExtract.csv.gz" "$CLUSTERDIR/bcPreclust.fastq.gz"
zcat "$CLUSTERDIR/clusters.csv.gz" | tail -n +2 | cut -f 4 -d, | sort -n | uniq -c > "$CLUSTERDIR/qc/clusters_distr.txt"
