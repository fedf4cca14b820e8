TAG=t13 NETS="resnet50 vgg16 googlenet squeezenet" PYTEST_EXTRA="--maxfail=12" bash tools/gpu_visit_r2.sh
TAG=a2 bash tools/gpu_aux_r2.sh
