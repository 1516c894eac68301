build/sort_lab 28 > gpurun_out/sort_lab2.txt 2>&1
cat gpurun_out/sort_lab2.txt
