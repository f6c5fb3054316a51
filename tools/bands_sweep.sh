for b in 2 3 4 5 6 8; do echo "cfg3 bands $b"; PICHA_B200_BANDS=$b WORKLOADS=cfg3 tools/ab.sh default; done
for b in 1 2 3 4; do echo "cfg5 bands $b"; PICHA_B200_BANDS=$b WORKLOADS=cfg5 tools/ab.sh default; done
