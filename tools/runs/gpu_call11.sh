#!/bin/bash
timeout 300 python tools/prof_tc.py 2
