// TEST INFRASTRUCTURE ONLY: forwards to the nlohmann/json 3.11.3 single header that ships inside the image's
// python environment (the reference expects it at open_spiel/json/, cloned by its install.sh; no network here).
#include "/opt/prime-rl/.venv/lib/python3.12/site-packages/include/cudnn_frontend/thirdparty/nlohmann/json.hpp"
