#!/bin/bash
timeout 120 python tools/run_once.py C5 148 1 2>&1 | tail -4
