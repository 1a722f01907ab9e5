#!/bin/bash
# How many host threads really run in parallel on this box (cgroup quota vs. logical CPUs)?  Times a fixed spin
# loop per thread for T = 1 .. nproc threads: flat times = real cores, growing times = oversubscribed quota.
set -e
cd "$(dirname "$0")"
gcc -O1 cpu_scaling_probe.c -o /tmp/cpu_scaling_probe -lpthread
echo "nproc=$(nproc) cpu.max=$(cat /sys/fs/cgroup/cpu.max 2>/dev/null || echo n/a) cfs_quota=$(cat /sys/fs/cgroup/cpu/cpu.cfs_quota_us 2>/dev/null || echo n/a)"
for t in 1 2 4 8 16 32 64 128; do /tmp/cpu_scaling_probe $t; done
