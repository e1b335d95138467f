// Compiles and links the C++ host mirror against libzkpor_b200.so and checks the reference's error behaviour that
// does not need a GPU (constructor guards, PaddingAccountAssets rule, "no CPU fallback").  Run by tests/test_abi.py.
#include <cstdio>
#include <cstring>
#include "zkpor_b200.hpp"

int main() {
    using namespace zkpor;
    // PaddingAccountAssets: gaps take the lowest unused indices first (src/utils/utils.go:147-186)
    std::vector<utils::AccountAsset> assets = {{3, 1, 2, 3, 4, 5}, {7, 9, 9, 9, 9, 9}};
    auto flat = utils::PaddingAccountAssets(assets);
    if (flat.size() != 300 || flat[0] != 0 || flat[6] != 1 || flat[18] != 3 || flat[19] != 1 || flat[24] != 4) { std::puts("FAIL padding"); return 1; }
    if (utils::GetAssetsCountOfUser(51) != 500) { std::puts("FAIL tier"); return 1; }
    int32_t ndev = 0; zkpor_device_count(&ndev);
    if (ndev == 0) {
        try { Context c(0); std::puts("FAIL: context without a GPU"); return 1; }
        catch (const Error &e) { if (!std::strstr(e.what(), "no CPU fallback")) { std::puts("FAIL message"); return 1; } }
        std::puts("OK (no GPU: construction refused loudly)");
        return 0;
    }
    Context ctx(0);
    Hash nil{};
    for (int bad = 0; bad < 3; bad++) {
        try {
            if (bad == 0) merkletree::FixedDepthMerkleTree t(ctx, 33, nil, 1);
            if (bad == 1) merkletree::FixedDepthMerkleTree t(ctx, 0, nil, 1);
            if (bad == 2) merkletree::FixedDepthMerkleTree t(ctx, 3, nil, 9);
            std::puts("FAIL guard"); return 1;
        } catch (const Error &) {}
    }
    merkletree::FixedDepthMerkleTree t(ctx, 8, nil, 100);
    Hash leaf{}; leaf[31] = 7;
    t.Set(5, leaf); t.Build();
    auto proof = t.GetProof(5);
    if (!merkletree::VerifyProof(ctx, t.Root(), 5, proof, leaf, 8)) { std::puts("FAIL verify"); return 1; }
    if (merkletree::VerifyProof(ctx, t.Root(), 4, proof, leaf, 8)) { std::puts("FAIL verify neg"); return 1; }
    // the batch loop refuses accounts that do not fill whole batches (the service pads the last one, src/witness/main.go:71-83)
    try { zkpor_cex_desc cex{}; witness::RunBatches(ctx, cex, nil, {}, {1, 2, 3}, 50, 2); std::puts("FAIL batches guard"); return 1; }
    catch (const Error &) {}
    std::puts("OK (GPU: tree build + proof verify through the C++ mirror)");
    return 0;
}
