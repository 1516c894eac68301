build/sort_lab 28 > gpurun_out/sort_lab4.txt 2>&1
cat gpurun_out/sort_lab4.txt
