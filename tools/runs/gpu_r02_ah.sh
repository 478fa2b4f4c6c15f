#!/bin/bash
timeout 120 python tools/time_launch_cpu.py
