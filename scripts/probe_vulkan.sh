#!/bin/bash
# Probe of the GPU box for what the reference's own `pt` would need: Vulkan loader + headers, an ICD with ray queries
# (lavapipe or the NVIDIA driver's), meson, a slang compiler.
echo "== tools"; for t in vulkaninfo meson slangc glslangValidator ninja cmake clang; do printf "%s: " $t; command -v $t || echo absent; done
echo "== loader / ICDs"; ls /usr/lib/x86_64-linux-gnu/libvulkan* /usr/lib/x86_64-linux-gnu/libvulkan_lvp* 2>&1 | head -5
ls /usr/share/vulkan/icd.d /etc/vulkan/icd.d 2>&1 | head -10
find / -xdev \( -name "vulkan.h" -o -name "vulkan.hpp" -o -name "libslang*" -o -name "nvidia_icd.json" -o -name "lvp_icd*.json" \) 2>/dev/null | head -10
echo "== driver"; nvidia-smi --query-gpu=name,driver_version --format=csv,noheader; ls /usr/lib/x86_64-linux-gnu/libnvidia-rtcore* /usr/lib/x86_64-linux-gnu/libGLX_nvidia* 2>&1 | head -4
echo "== cpu"; nproc; grep -m1 "model name" /proc/cpuinfo
