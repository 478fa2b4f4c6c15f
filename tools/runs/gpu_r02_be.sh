#!/bin/bash
timeout 300 python tools/check_tc4.py 2>&1 | tail -30
