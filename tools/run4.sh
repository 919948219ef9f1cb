cd spla_b200/lib
echo "=== test_cuda_backend 12"; timeout 300 ./test_cuda_backend 12 > /tmp/o.txt 2>&1; echo "rc=$?"; tail -40 /tmp/o.txt
