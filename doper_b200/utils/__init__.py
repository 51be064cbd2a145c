"""Mirror of the reference ``doper/utils`` pieces the rollout path's callers need (scene ingest)."""
