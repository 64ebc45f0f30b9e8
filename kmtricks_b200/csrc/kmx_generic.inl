// kmx_generic.inl -- generic (sort-based) stage 2 and sparse-row stage 3 drivers (included by kmx_api.cu)
static int count_generic(kmx_ctx* ctx, uint32_t sample, uint32_t hard_min)
{
  (void)sample; (void)hard_min;
  return fail(ctx, KMX_ERR_ARG, "generic count path not built yet");
}
static int merge_sparse(kmx_ctx* ctx, uint32_t partition, const kmx_merge_params* mp, kmx_merge_result* res,
                        const std::vector<MergeList>& hl, u64 max_n, u64 tot_n)
{
  (void)partition; (void)mp; (void)res; (void)hl; (void)max_n; (void)tot_n;
  return fail(ctx, KMX_ERR_ARG, "sparse merge path not built yet");
}
