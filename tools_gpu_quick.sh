#!/bin/bash
timeout 120 python tools/conv_bench.py 512,24576 conv1_2,conv2_1,conv2_2,conv3_2 2>&1 | tail -10
