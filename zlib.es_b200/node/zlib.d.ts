// Same declarations as the reference's dist/tsc/zlib.d.ts, plus the batch forms.
export declare function inflate(input: Uint8Array): Uint8Array;
export declare function deflate(input: Uint8Array): Uint8Array;
export declare function inflateBatch(inputs: Uint8Array[]): Uint8Array[];
export declare function deflateBatch(inputs: Uint8Array[]): Uint8Array[];
